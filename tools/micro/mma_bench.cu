// Microbenchmark: tensor-pipe throughput of the two candidate mod-p contraction paths for the dense
// echelon's trailing update (BASELINE north star: "FP64 DMMA with the exact-accumulation bound on
// k-blocks, or int8 IMMA on limb-split operands"), through the warp-level mma.sync forms that compile
// for sm_100a:
//   mode 0  mma.sync.m8n8k4   f64  (DMMA)       256 MAC / instruction
//   mode 1  mma.sync.m16n8k32 u8 x u8 -> s32   4096 MAC / instruction
//   mode 2  IMAD.WIDE u64 accumulation on the CUDA cores (baseline without tensor cores)
// Register-resident operands, ILP independent accumulator chains per warp.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8;

__global__ void dmma_bench(double* out, int iters) {
    double a = threadIdx.x * 0.5, b = blockIdx.x + 1.0;
    double c[ILP][2];
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void imma_bench(int* out, int iters) {
    unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, blockIdx.x}, b[2] = {threadIdx.x * 5u, 11u};
    int c[ILP][4];
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void imad_bench(unsigned long long* out, int iters) {
    unsigned a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7;
    unsigned long long c[ILP];
    for (int i = 0; i < ILP; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { c[i] += (unsigned long long)a * b; a += (unsigned)c[i]; }
    }
    unsigned long long s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    void* d; cudaMalloc(&d, 64 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int mode = 0; mode < 3; ++mode) {
            const int ctas = 148 * 2, thr = warps * 16;
            auto launch = [&]() {
                if (mode == 0) dmma_bench<<<ctas, thr>>>((double*)d, iters);
                else if (mode == 1) imma_bench<<<ctas, thr>>>((int*)d, iters);
                else imad_bench<<<ctas, thr>>>((unsigned long long*)d, iters);
            };
            launch();
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double per = mode == 0 ? 256.0 : mode == 1 ? 4096.0 : 32.0;
            const double mac = (double)ctas * (thr / 32) * iters * ILP * per;
            printf("warps/SM=%2d mode=%d (%s)  %.3f ms  %.2f TMAC/s  err=%s\n", warps, mode,
                   mode == 0 ? "DMMA m8n8k4 f64" : mode == 1 ? "IMMA m16n8k32 u8" : "IMAD.WIDE u64", ms, mac / ms / 1e9,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
