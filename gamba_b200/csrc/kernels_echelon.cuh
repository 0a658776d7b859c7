// Echelon of the dense D' block (role of the reference's closed
// kernel/echelon_u16|u32 translation units, CMakeLists.txt:81-90) and the
// dense -> sparse compaction of the new rows (open reference code:
// source/matrix.hpp:820-880).
//
// Exact Gauss-Jordan as ONE cooperative kernel, one grid barrier per pivot (a variant clearing 8
// pivots per pass against a snapshotted panel was bit-exact but 1.8x slower on cyclic9-15):
// every CTA re-derives the next pivot (left-most lead among the remaining
// rows) from the double-buffered lead table, so selection needs no barrier of
// its own; the pivot column is never written during an update (it is "dead"
// for all rows but its pivot row afterwards), which removes the barrier between
// reading the multipliers and updating the rows. Pivots come out in ascending
// column order and every pivot column is cleared from ALL other rows, so the
// result is the canonical reduced row echelon form that the host driver (and
// basis.hpp:725-727 in the reference) expects, rows sorted by ascending pivot.
#pragma once
#include <cooperative_groups.h>

#include "layout.cuh"

namespace gcu {
namespace cg = cooperative_groups;

constexpr int ECH_THREADS = 512;
constexpr uint32_t ECH_INF = 0xffffffffu;

struct EchArgs {
    Fp f;
    uint32_t n_rows, n_cols, ld;    // D' is n_rows x n_cols (non-pivot columns), row-major u32, row pitch ld
    uint32_t* d;
    const uint32_t* row_nz;         // [n_rows]
    uint32_t* lead;                 // [2][n_rows]
    uint32_t* inv_lead;             // [n_rows] inverse of the row's lead value
    uint32_t* nz_list;              // [n_rows] rows that are not identically zero
    uint32_t* piv_row;              // [n_rows]
    uint32_t* piv_col;              // [n_rows]
    uint8_t* is_piv_col;            // [n_cols]
    uint32_t* counts;               // [0] n_nz, [1] n_piv
    uint64_t* out_ptr;              // [n_rows + 1] CSR row pointers of the result
};

__device__ __forceinline__ uint32_t block_min_u32(uint32_t v, uint32_t* sh) {
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : ECH_INF;
        for (int o = 16; o; o >>= 1) w = min(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (threadIdx.x == 0) sh[0] = w;
    }
    __syncthreads();
    v = sh[0];
    return v;
}

__device__ __forceinline__ uint64_t block_min_u64(uint64_t v, uint64_t* sh) {
    for (int o = 16; o; o >>= 1) { uint64_t w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint64_t w = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : ~0ull;
        for (int o = 16; o; o >>= 1) { uint64_t z = __shfl_xor_sync(0xffffffffu, w, o); w = z < w ? z : w; }
        if (threadIdx.x == 0) sh[0] = w;
    }
    __syncthreads();
    v = sh[0];
    return v;
}

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t* sh) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0u;
        for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) sh[0] = w;
    }
    __syncthreads();
    v = sh[0];
    return v;
}


struct EmitArgs {
    Fp f;
    uint32_t n_cols, n_piv, ld;
    const uint32_t* d;
    const uint32_t* piv_row;
    const uint32_t* piv_col;
    const uint8_t* is_piv_col;
    const uint64_t* out_ptr;
    uint32_t* out_col;
    uint32_t* out_val;              // standard or Montgomery residues per f.mont_bits
};

// One CTA per new row: normalise (make monic), drop dead columns, compact in order.
__global__ void __launch_bounds__(256) emit_kernel(EmitArgs a) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    const uint32_t s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* row = a.d + (uint64_t)a.piv_row[s] * a.ld;
    const uint32_t c = a.piv_col[s];
    __shared__ uint32_t s_scale;
    if (tid == 0) {
        uint32_t inv = fp_inv32(row[c], a.f.p);
        if (a.f.mont_bits) inv = fp_mul(inv, a.f.r_mod_p, a.f);
        s_scale = inv; s_base = 0;
    }
    __syncthreads();
    const uint32_t scale = s_scale;
    const uint64_t o = a.out_ptr[s];
    for (uint32_t j0 = c; j0 < a.n_cols; j0 += 256) {
        const uint32_t j = j0 + tid;
        uint32_t v = 0;
        if (j < a.n_cols && (j == c || !a.is_piv_col[j])) v = row[j];
        const uint32_t bal = __ballot_sync(0xffffffffu, v != 0);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint32_t pre = s_base;
        for (uint32_t w = 0; w < warp; ++w) pre += s_warp[w];
        if (v) {
            const uint64_t pos = o + pre + __popc(bal & ((1u << lane) - 1u));
            a.out_col[pos] = j;
            a.out_val[pos] = fp_mul(v, scale, a.f);
        }
        __syncthreads();
        if (tid == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
}

}  // namespace gcu
