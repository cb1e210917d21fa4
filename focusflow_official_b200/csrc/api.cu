// libffcorr: error buffer, version and device info.
#include <stdarg.h>

#include "common.cuh"

namespace ffcorr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMsB200;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = kNumSMsB200;
        cached_dev = dev;
        cached_sms = n;
    }
    return cached_sms;
}

int check_levels(int num_levels, int h, int w, const char* who) {
    FFCORR_REQUIRE(num_levels >= 1 && num_levels <= FFCORR_MAX_LEVELS, FFCORR_EINVAL,
                   "%s: num_levels=%d outside [1,%d]", who, num_levels, FFCORR_MAX_LEVELS);
    FFCORR_REQUIRE(h >= 1 && w >= 1 && h <= 16384 && w <= 16384, FFCORR_EINVAL,
                   "%s: h=%d w=%d outside [1,16384]", who, h, w);
    FFCORR_REQUIRE((h >> (num_levels - 1)) >= 1 && (w >> (num_levels - 1)) >= 1, FFCORR_EINVAL,
                   "%s: %dx%d map is too small for %d pyramid levels", who, h, w, num_levels);
    return FFCORR_OK;
}

}  // namespace ffcorr

extern "C" int ffcorr_version(void) { return FFCORR_VERSION; }

extern "C" const char* ffcorr_last_error(void) { return ffcorr::g_err; }

extern "C" int ffcorr_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    FFCORR_CUDA(cudaGetDevice(&dev));
    int sms = 0, maj = 0, min = 0;
    FFCORR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FFCORR_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    FFCORR_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    return FFCORR_OK;
}

extern "C" int ffcorr_set_l2_fetch_granularity(int bytes) {
    FFCORR_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, FFCORR_EINVAL, "l2 fetch granularity must be 32, 64 or 128");
    FFCORR_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
    return FFCORR_OK;
}

extern "C" int ffcorr_get_l2_fetch_granularity(int* bytes) {
    size_t v = 0;
    FFCORR_CUDA(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
    if (bytes) *bytes = (int)v;
    return FFCORR_OK;
}
