// Micro-benchmark (not part of the product): random-gather capacity of the B200 memory system for the
// request shapes the lookup could use.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_gather mb_gather.cu
//   mode 0: LDG.128, every lane its own random 64-byte granule (16 B used)        -> 1 sector / lane
//   mode 1: LDG.128, 4 lanes cover one random 64-byte granule                     -> full 64 B tile per 4 lanes
//   mode 2: LDG.128, 8 lanes cover one random 128-byte line
//   mode 3: LDGSTS 16 B, every lane its own granule (like the row-streaming lookup)
//   mode 4: LDGSTS 16 B, 4 lanes per granule
//   mode 5: cp.async.bulk 64 B per lane (TMA engine), mbarrier completion
//   mode 6: LDG.128 x4 per lane: a lane reads its whole random 64-byte granule
//   mode 7: LDG.128, 16 lanes cover one random aligned 256-byte block
//   mode 8: LDG.128, 32 lanes cover one random aligned 512-byte block
//   mode 9: LDG.128, 8 lanes cover 128 bytes starting at a random 64-byte granule (half the runs straddle two lines)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int MODE>
__global__ void __launch_bounds__(256) gather_kernel(const float4* __restrict__ buf, uint32_t granules, int iters, float* sink) {
    __shared__ __align__(128) float4 stage[256 * 4];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t gtid = blockIdx.x * 256 + tid;
    float acc = 0.f;
    if (MODE == 5) {
        if (tid == 0) { asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar))); }
        __syncthreads();
    }
    uint32_t phase = 0;
    for (int it = 0; it < iters; it += 8) {
        if (MODE <= 2 || MODE >= 6) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint32_t key = (MODE == 0 || MODE == 6) ? gtid : (MODE == 1 ? (gtid >> 2) : (MODE == 7 ? (gtid >> 4) : (MODE == 8 ? (gtid >> 5) : (gtid >> 3))));
                uint32_t g = hash32(key * 977u + (uint32_t)(it + u) * 0x9e3779b9u) % granules;
                size_t idx = (size_t)g * 4 + (MODE == 1 ? (lane & 3) : 0);
                if (MODE == 2) idx = ((size_t)(g & ~1u)) * 4 + (lane & 7);
                if (MODE == 7) idx = ((size_t)(g & ~3u)) * 4 + (lane & 15);
                if (MODE == 8) idx = ((size_t)(g & ~7u)) * 4 + (lane & 31);
                if (MODE == 9) idx = (size_t)(g < granules - 2 ? g : 0) * 4 + (lane & 7);
                v[u] = __ldg(buf + idx);
                if (MODE == 6) {
                    float4 b = __ldg(buf + idx + 1), c = __ldg(buf + idx + 2), d = __ldg(buf + idx + 3);
                    v[u].x += b.x + c.y + d.z;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].w;
        } else if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t key = (MODE == 3) ? gtid : (gtid >> 2);
                uint32_t g = hash32(key * 977u + (uint32_t)(it + u) * 0x9e3779b9u) % granules;
                size_t idx = (size_t)g * 4 + (MODE == 4 ? (lane & 3) : 0);
                uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[u * 256 + tid]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(buf + idx) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            acc += stage[tid].x;
            it -= 4;   // 4 requests per trip
        } else if (MODE == 5) {
            // every lane of warp 0..7 issues one 64-byte bulk copy per trip into its own slot
            const uint32_t barp = (uint32_t)__cvta_generic_to_shared(&bar);
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(barp), "r"(256 * 64) : "memory");
            __syncthreads();
            uint32_t g = hash32(gtid * 977u + (uint32_t)it * 0x9e3779b9u) % granules;
            uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[tid * 4]);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(dst),
                         "l"(buf + (size_t)g * 4), "r"(barp) : "memory");
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(barp), "r"(phase) : "memory");
            }
            phase ^= 1;
            acc += stage[tid * 4].x;
            __syncthreads();
            it -= 7;   // 1 request per trip
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int MODE>
void run(const float4* buf, uint32_t granules, float* sink, int blocks, int iters, const char* label, double lanes_per_req, double bytes_per_req) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_kernel<MODE><<<blocks, 256>>>(buf, granules, iters, sink);
    cudaEventRecord(e0);
    gather_kernel<MODE><<<blocks, 256>>>(buf, granules, iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    double reqs = (double)blocks * 256 * iters / lanes_per_req;
    printf("%-52s %8.3f ms  %7.2f G req/s  %7.0f GB/s (at %g B/req)  %s\n", label, ms, reqs / ms / 1e6, reqs * bytes_per_req / ms / 1e6,
           bytes_per_req, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)4 << 30;   // 4 GB: far beyond L2
    float4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    float* sink; cudaMalloc(&sink, 4);
    const uint32_t granules = (uint32_t)(bytes / 64);
    const int blocks = 148 * 8 * 4;
    const int iters = 64;
    run<0>(buf, granules, sink, blocks, iters, "LDG.128  1 lane / 64B granule (16 B used)", 1, 64);
    run<1>(buf, granules, sink, blocks, iters, "LDG.128  4 lanes / 64B granule", 4, 64);
    run<2>(buf, granules, sink, blocks, iters, "LDG.128  8 lanes / 128B line", 8, 128);
    run<6>(buf, granules, sink, blocks, iters, "LDG.128x4 1 lane reads whole 64B granule", 1, 64);
    run<7>(buf, granules, sink, blocks, iters, "LDG.128 16 lanes / aligned 256B block", 16, 256);
    run<8>(buf, granules, sink, blocks, iters, "LDG.128 32 lanes / aligned 512B block", 32, 512);
    run<9>(buf, granules, sink, blocks, iters, "LDG.128  8 lanes / 128B at a random 64B offset", 8, 128);
    run<3>(buf, granules, sink, blocks, iters, "LDGSTS.16 1 lane / 64B granule", 1, 64);
    run<4>(buf, granules, sink, blocks, iters, "LDGSTS.16 4 lanes / 64B granule", 4, 64);
    run<5>(buf, granules, sink, blocks, iters, "cp.async.bulk 64 B / lane", 1, 64);
    return 0;
}
