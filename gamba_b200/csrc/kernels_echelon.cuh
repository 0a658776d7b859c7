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

constexpr int ECH_GROUP = 128;                       // threads that update one row together
constexpr int ECH_GROUPS = ECH_THREADS / ECH_GROUP;

__device__ __forceinline__ void group_bar(uint32_t grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(ECH_GROUP) : "memory");
}

__global__ void __launch_bounds__(ECH_THREADS, 1) echelon_kernel(EchArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint64_t sh64[32];
    __shared__ uint32_t sh32[32];
    __shared__ uint32_t sh_grp[ECH_GROUPS][4];
    const uint32_t tid = threadIdx.x, nb = gridDim.x, bid = blockIdx.x;
    const uint32_t grp = tid / ECH_GROUP, gt = tid % ECH_GROUP, gw = gt >> 5;
    const uint32_t nr = a.n_rows, nc = a.n_cols, p = a.f.p;
    uint32_t* lead0 = a.lead;
    uint32_t* lead1 = a.lead + nr;

    // leads of the rows (block per row), the inverse of every lead value, the list of non-zero rows
    for (uint32_t r = bid; r < nr; r += nb) {
        uint32_t l = ECH_INF;
        if (a.row_nz[r]) {
            const uint32_t* row = a.d + (uint64_t)r * a.ld;
            for (uint32_t j0 = 0; j0 < nc && l == ECH_INF; j0 += ECH_THREADS) {
                const uint32_t j = j0 + tid;
                l = block_min_u32(j < nc && row[j] ? j : ECH_INF, sh32);
            }
            if (tid == 0 && l != ECH_INF) a.inv_lead[r] = fp_inv32(row[l], p);
        }
        if (tid == 0) lead0[r] = l;
    }
    for (uint32_t j = bid * ECH_THREADS + tid; j < nc; j += nb * ECH_THREADS) a.is_piv_col[j] = 0;
    grid.sync();
    if (bid == 0 && tid == 0) {
        uint32_t n = 0;
        for (uint32_t r = 0; r < nr; ++r) if (lead0[r] != ECH_INF) a.nz_list[n++] = r;
        a.counts[0] = n;
    }
    grid.sync();
    const uint32_t nnz_rows = a.counts[0];

    uint32_t step = 0;
    for (;; ++step) {
        const uint32_t* lc = (step & 1) ? lead1 : lead0;
        uint32_t* ln = (step & 1) ? lead0 : lead1;
        // every CTA derives the same pivot: smallest lead, ties to the first row
        uint64_t key = ~0ull;
        for (uint32_t i = tid; i < nnz_rows; i += ECH_THREADS) {
            const uint32_t l = lc[a.nz_list[i]];
            if (l != ECH_INF) { const uint64_t k = ((uint64_t)l << 32) | i; key = k < key ? k : key; }
        }
        key = block_min_u64(key, sh64);
        if (key == ~0ull) break;
        const uint32_t c = (uint32_t)(key >> 32);
        const uint32_t pr = a.nz_list[(uint32_t)key];
        const uint32_t* prow = a.d + (uint64_t)pr * a.ld;
        const uint32_t inv = a.inv_lead[pr];
        if (bid == 0 && tid == 0) { a.piv_row[step] = pr; a.piv_col[step] = c; a.is_piv_col[c] = 1; }
        // rows are dealt to groups of 128 threads: 4 rows per CTA in flight
        for (uint32_t i = bid * ECH_GROUPS + grp; i < nnz_rows; i += nb * ECH_GROUPS) {
            const uint32_t r = a.nz_list[i];
            if (r == pr) { if (gt == 0) ln[r] = ECH_INF; continue; }
            uint32_t* row = a.d + (uint64_t)r * a.ld;
            const uint32_t fv = row[c];
            const uint32_t lr = lc[r];
            if (fv == 0) { if (gt == 0) ln[r] = lr; continue; }
            const uint32_t mult = p - fp_mul(fv, inv, a.f);
            uint32_t nl = ECH_INF;
            for (uint32_t j0 = c + 1; j0 < nc; j0 += 4 * ECH_GROUP) {
                uint32_t pv[4], v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t j = j0 + u * ECH_GROUP + gt;
                    pv[u] = j < nc ? prow[j] : 0u;
                    v[u] = j < nc ? row[j] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t j = j0 + u * ECH_GROUP + gt;
                    if (pv[u]) { v[u] = fp_reduce64((uint64_t)v[u] + (uint64_t)mult * pv[u], a.f); row[j] = v[u]; }
                    if (v[u] && nl == ECH_INF) nl = j;
                }
            }
            if (lr != ECH_INF) {            // a remaining row: its lead was c, find the new one (and its inverse)
                for (int o = 16; o; o >>= 1) nl = min(nl, __shfl_xor_sync(0xffffffffu, nl, o));
                if ((gt & 31) == 0) sh_grp[grp][gw] = nl;
                group_bar(grp);             // also orders the row[] stores before the read below
                if (gt == 0) {
                    nl = min(min(sh_grp[grp][0], sh_grp[grp][1]), min(sh_grp[grp][2], sh_grp[grp][3]));
                    ln[r] = nl;
                    if (nl != ECH_INF) a.inv_lead[r] = fp_inv32(row[nl], p);
                }
                group_bar(grp);
            } else if (gt == 0) ln[r] = ECH_INF;
        }
        grid.sync();
    }
    const uint32_t npiv = step;
    if (bid == 0 && tid == 0) a.counts[1] = npiv;

    // non-zeros per pivot row over the live columns (its own pivot + non-pivot columns)
    for (uint32_t s = bid; s < npiv; s += nb) {
        const uint32_t* row = a.d + (uint64_t)a.piv_row[s] * a.ld;
        const uint32_t c = a.piv_col[s];
        uint32_t n = 0;
        for (uint32_t j = c + tid; j < nc; j += ECH_THREADS) n += (j == c || (!a.is_piv_col[j] && row[j])) ? 1u : 0u;
        n = block_sum_u32(n, sh32);
        if (tid == 0) a.out_ptr[s + 1] = n;
    }
    grid.sync();
    if (bid == 0 && tid == 0) {
        uint64_t acc = 0; a.out_ptr[0] = 0;
        for (uint32_t s = 0; s < npiv; ++s) { acc += a.out_ptr[s + 1]; a.out_ptr[s + 1] = acc; }
    }
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
