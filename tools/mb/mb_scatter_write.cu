// Micro-benchmark (not part of the product): how fast can the B200 memory system absorb the WRITE PATTERN of
// the correlation-volume GEMM, independent of the GEMM itself?  148 persistent CTAs x 4 warps; CTA handles
// tile t = (batch, 128-query block, 256-column n-tile), n fastest, warp w owns 32 queries; every query gets
// `piece` contiguous bytes per chunk, queries are `map_bytes` apart.
//   mode 0: linear streaming write of the same number of bytes (upper bound)
//   mode 1: unfused tiled layout: n-tile = 1 KB contiguous per query (4 chunks x 256 B adjacent)
//   mode 2: fused layout: chunk c = tile row (4*sgy + c), 256 B at (ty*tw0 + 4*sgx)*64 B
//   mode 4: 8x32-pixel n-tiles: two 512 B tile-row pieces per query (chunks 0,2 -> row 0; 1,3 -> row 1)
//   mode 3: like 1 but m fastest (CTAs at one time write the same columns of 148 different query blocks)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_scatter_write mb_scatter_write.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct P { int B, N, tiles_m, tiles_n, sgw, tw0; long long map_floats; int mode; float *l1, *l2, *l3; int extra; };

__global__ void __launch_bounds__(128) wr(float* __restrict__ out, P p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_batch = p.tiles_m * p.tiles_n;
    const long long num_tiles = (long long)tiles_per_batch * p.B;
    const float4 val = make_float4(1.f, 2.f, 3.f, 4.f);
    for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int b = (int)(t / tiles_per_batch);
        const int r = (int)(t - (long long)b * tiles_per_batch);
        int mt, nt;
        if (p.mode == 3) { nt = r / p.tiles_m; mt = r % p.tiles_m; } else { mt = r / p.tiles_n; nt = r % p.tiles_n; }
        const int row0 = mt * 128 + warp * 32;
        if (p.mode == 0) {
            // same bytes, written linearly: tile t -> 128 KB contiguous
            float4* dst = reinterpret_cast<float4*>(out) + t * 8192 + warp * 2048;
            for (int i = 0; i < 64; ++i) dst[i * 32 + lane] = val;
            continue;
        }
        for (int c = 0; c < 4; ++c) {
            long long col;   // float offset inside the query map of this chunk's 64 floats
            if (p.extra && p.mode == 2) {
                // the other pyramid levels, written like the fused build does: L1 128 B per chunk pair, L2 32 B per
                // chunk pair, L3 8 B per chunk pair (tw1 = 20, tw2 = 10, tw3 = 6 tiles; maps 1920 / 480 / 192 floats)
                const int sgy = nt / p.sgw, sgx = nt - sgy * p.sgw;
                if (c & 1) {
                    for (int rr = 0; rr < 32; rr += 4) {
                        const int row = row0 + rr + (lane >> 3);
                        if (row < p.N && (p.extra & 1))
                            *reinterpret_cast<float4*>(p.l1 + ((long long)b * p.N + row) * 1920 + ((sgy * 2 + (c >> 1)) * 20 + sgx * 2) * 16 + (lane & 7) * 4) = val;
                    }
                    const int row = row0 + lane;
                    if (row < p.N && (p.extra & 2) && sgy < 3) {
                        float4* d2 = reinterpret_cast<float4*>(p.l2 + ((long long)b * p.N + row) * 480 + (sgy * 10 + sgx) * 16 + (c - 1) * 4);
                        d2[0] = val; d2[1] = val;
                    }
                    if (row < p.N && (p.extra & 4))
                        *reinterpret_cast<float2*>(p.l3 + ((long long)b * p.N + row) * 192 + ((sgy >> 1) * 6 + (sgx >> 1)) * 16 + ((sgy & 1) * 2 + (c >> 1)) * 4 + (sgx & 1) * 2) = make_float2(1.f, 2.f);
                }
            }
            if (p.mode == 4) {
                const int sy = nt / p.sgw, sx = nt - sy * p.sgw;      // sgw = n-tiles per row (32 px wide)
                col = ((long long)(sy * 2 + (c & 1)) * p.tw0 + sx * 8 + (c >> 1) * 4) * 16;
                if ((sx * 8 + (c >> 1) * 4 + 4) > p.tw0 + 1 || col + 64 > p.map_floats) continue;
            } else if (p.mode == 2) {
                const int sgy = nt / p.sgw, sgx = nt - sgy * p.sgw;
                col = ((long long)(sgy * 4 + c) * p.tw0 + sgx * 4) * 16;
                if (col + 64 > p.map_floats) continue;
            } else {
                col = (long long)nt * 256 + c * 64;
                if (col + 64 > p.map_floats) continue;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int rr = 0; rr < 32; rr += 4) {
                    const int row = row0 + rr + (lane >> 3);
                    if (row < p.N) {
                        float* d = out + ((long long)b * p.N + row) * p.map_floats + col + h * 32 + (lane & 7) * 4;
                        *reinterpret_cast<float4*>(d) = val;
                    }
                }
        }
    }
}

int main() {
    P p; p.extra = 0; p.B = 8; p.N = 7332; p.tiles_m = 58; p.tw0 = 39; p.map_floats = 7488;
    float* out; size_t bytes = (size_t)p.B * p.N * 7680 * 4 + (64 << 20);
    cudaMalloc(&out, bytes);
    cudaMemset(out, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[5] = {"linear", "unfused tiled (1 KB/query/tile)", "fused (4 x 256 B tile rows)", "unfused, m fastest",
                            "fused 8x32 px (2 x 512 B)"};
    for (int mode = 0; mode < 5; ++mode) {
        p.mode = mode;
        if (mode == 2) { p.sgw = 10; p.tiles_n = 30; } else if (mode == 4) { p.sgw = 5; p.tiles_n = 30; } else { p.sgw = 0; p.tiles_n = 30; }
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            wr<<<148, 128>>>(out, p);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gb = (double)p.B * p.N * 7488 * 4 / 1e9;
        if (mode == 0) gb = (double)p.B * p.tiles_m * p.tiles_n * 131072 / 1e9;
        printf("%-36s %7.3f ms  %7.0f GB/s  %s\n", names[mode], ms, gb / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    // aligned variants: tile rows padded to 40 tiles (every 256 B piece covers two whole 128 B lines)
    for (int mode = 1; mode <= 4; ++mode) {
        if (mode == 3) continue;
        p.mode = mode; p.tw0 = 40; p.map_floats = 7680;
        p.sgw = mode == 2 ? 10 : 5; p.tiles_n = 30;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            wr<<<148, 128>>>(out, p);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gb = (double)p.B * p.N * 7680 * 4 / 1e9;
        printf("tw0=40: %-28s %7.3f ms  %7.0f GB/s\n", names[mode], ms, gb / ms * 1e3);
    }
    {
        size_t n = (size_t)p.B * p.N;
        cudaMalloc(&p.l1, n * 1920 * 4); cudaMalloc(&p.l2, n * 480 * 4); cudaMalloc(&p.l3, n * 192 * 4);
        const char* en[4] = {"L0 only", "L0+L1", "L0+L1+L2", "L0+L1+L2+L3"};
        const int ex[4] = {0, 1, 3, 7};
        const double fl[4] = {7680, 7680 + 1920, 7680 + 1920 + 480, 7680 + 1920 + 480 + 160};
        for (int k = 0; k < 4; ++k) {
            p.mode = 2; p.tw0 = 40; p.map_floats = 7680; p.sgw = 10; p.tiles_n = 30; p.extra = ex[k];
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                wr<<<148, 128>>>(out, p);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double gb = (double)p.B * p.N * fl[k] * 4 / 1e9;
            printf("fused pattern %-14s            %7.3f ms  %7.0f GB/s  %s\n", en[k], ms, gb / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
        }
        p.extra = 0;
    }
    p.tw0 = 39; p.map_floats = 7488;
    // more CTAs per SM for the linear case (LSU-issue bound with 4 warps/SM?)
    for (int mult = 2; mult <= 8; mult *= 2) {
        p.mode = 1; p.tiles_n = 30;
        cudaEventRecord(e0);
        wr<<<148 * mult, 128>>>(out, p);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gb = (double)p.B * p.N * 7488 * 4 / 1e9;
        printf("unfused tiled, %d CTAs/SM               %7.3f ms  %7.0f GB/s\n", mult, ms, gb / ms * 1e3);
    }
    return 0;
}
