// libffcorr: error buffer, version and device info.
#include <stdarg.h>

#include <cudaTypedefs.h>

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

static PFN_cuTensorMapEncodeTiled resolve_encode_fn() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    return nullptr;
}

static PFN_cuTensorMapEncodeTiled get_encode_fn() {
    // C++11 magic static: initialised exactly once, concurrent first callers (PyTorch's autograd worker threads reach
    // this from the backward) block until it is
    static const PFN_cuTensorMapEncodeTiled fn = resolve_encode_fn();
    return fn;
}

int encode_tensor_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle, const char* what) {
    PFN_cuTensorMapEncodeTiled fn = get_encode_fn();
    FFCORR_REQUIRE(fn != nullptr, FFCORR_EDEVICE, "cuTensorMapEncodeTiled is not available from the driver");
    // The encoder is a DRIVER call and needs a current context.  A thread that has only been handed a device by
    // the runtime (PyTorch's autograd worker running a backward as its first CUDA work) has none bound yet, and the
    // call fails with CUDA_ERROR_INVALID_CONTEXT; cudaFree(nullptr) binds the primary context (once per thread).
    static thread_local bool context_bound = false;
    if (!context_bound) {
        FFCORR_CUDA(cudaFree(nullptr));
        context_bound = true;
    }
    FFCORR_REQUIRE(rank >= 1 && rank <= 5, FFCORR_EINVAL, "tensor map rank %d", rank);
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i + 1 < rank) st[i] = strides_bytes[i];
    }
    CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FFCORR_REQUIRE(r == CUDA_SUCCESS, FFCORR_EINVAL, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return FFCORR_OK;
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
