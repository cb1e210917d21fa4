// Shared helpers for libffcorr (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ffcorr.h"

namespace ffcorr {

// thread-local last-error buffer (api.cu)
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return FFCORR_OK;
}

#define FFCORR_REQUIRE(cond, code, ...)          \
    do {                                         \
        if (!(cond)) {                           \
            ::ffcorr::set_error(__VA_ARGS__);    \
            return (code);                       \
        }                                        \
    } while (0)

#define FFCORR_CUDA(expr)                                                        \
    do {                                                                         \
        cudaError_t e_ = (expr);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            ::ffcorr::set_error("%s: %s", #expr, cudaGetErrorString(e_));        \
            return (int)e_;                                                      \
        }                                                                        \
    } while (0)

constexpr int kNumSMsB200 = 148;

int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device

// shared argument validation for the pyramid-shaped entry points (api.cu)
int check_levels(int num_levels, int h, int w, const char* who);

// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no libcuda link); rank <= 5, no interleave,
// zero OOB fill.  dims/box are innermost-first, strides_bytes has rank-1 entries (dims 1..rank-1).
int encode_tensor_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle, const char* what);

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Tiled ("T4") level geometry: 4x4-pixel tiles of 64 bytes, row-major over [tiled_th][tiled_tw].  The tile-row
// pitch is kept EVEN so every tile row starts on a 128-byte line: the fused build writes 256-byte (level 0) and
// 128-byte (level 1) pieces per query, and pieces that straddle lines cost 15 % of the achievable write
// bandwidth (tools/mb/mb_scatter_write.cu: 4.8 vs 5.7 TB/s).
__host__ __device__ inline int tiled_th(int h_level) { return ceil_div(h_level, 4); }
__host__ __device__ inline int tiled_tw(int w_level) { return (ceil_div(w_level, 4) + 1) & ~1; }

}  // namespace ffcorr
