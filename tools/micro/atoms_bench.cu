// Microbenchmark: shared-memory update throughput per SM for the accumulator tile.
//  mode 0: plain 64-bit read-modify-write (LDS.64 / STS.64), random columns, one warp = one row
//  mode 1: native 32-bit ATOMS.ADD without return, random columns
//  mode 2: two 32-bit ATOMS.ADD (split lo/hi accumulators)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void bench(unsigned long long* out, int iters, int tile) {
    extern __shared__ unsigned long long sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* acc = sm + (size_t)warp * tile;
    unsigned* acc32 = reinterpret_cast<unsigned*>(acc);
    for (int i = lane; i < tile; i += 32) acc[i] = 0;
    __syncwarp();
    unsigned x = 1234567u * (threadIdx.x + 1) + blockIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        const unsigned c = (x >> 8) % (unsigned)tile;
        const unsigned v = x & 0xffff;
        if (MODE == 0) { acc[c] += (unsigned long long)v * 31337u; __syncwarp(); }
        if (MODE == 1) { atomicAdd(&acc32[c], v); }
        if (MODE == 2) { atomicAdd(&acc32[2 * c], v); atomicAdd(&acc32[2 * c + 1], v >> 3); }
    }
    __syncwarp();
    long long t1 = clock64();
    unsigned long long s = 0;
    for (int i = lane; i < tile; i += 32) s += acc[i];
    if (lane == 0) out[blockIdx.x * (blockDim.x >> 5) + warp] = (unsigned long long)(t1 - t0) ^ (s & 1);
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 1 << 20);
    const int tile = 3584, iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int mode = 0; mode < 3; ++mode) {
            const int ctas = 148 * 2, thr = warps * 16;  // 2 CTAs per SM
            size_t smem = (size_t)(thr / 32) * tile * 8;
            auto k = mode == 0 ? bench<0> : mode == 1 ? bench<1> : bench<2>;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<ctas, thr, smem>>>(d, iters, tile);
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            k<<<ctas, thr, smem>>>(d, iters, tile);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            double upd = (double)ctas * thr * iters;
            printf("warps/SM=%2d mode=%d  %.3f ms  %.1f Gupd/s  (%.2f updates/clk/SM at 1.9GHz) err=%s\n", warps, mode, ms,
                   upd / ms / 1e6, upd / ms / 1e-3 / 148 / 1.9e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
