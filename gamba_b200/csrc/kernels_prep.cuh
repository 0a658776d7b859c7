// Staging kernels: turn the block-form CSR handed over at the ABI (column lists +
// shared coefficient arrays) into the bucketed device layout of
// layout.cuh. One warp per row; coefficients are gathered from the shared arrays
// exactly once per step (rows share arrays ~25:1, source/matrix.hpp:189-193).
#pragma once
#include "layout.cuh"

namespace gcu {

struct PrepArgs {
    Layout L;
    Fp f;
    const uint64_t* row_ptr;
    const uint32_t* col;
    const uint32_t* coeff_id;
    const uint64_t* coeff_off;
    const uint32_t* pool;
    uint32_t* seg;        // in: zero / out of count: per-bin counts / after scan: starts
    uint2* ent;
    uint32_t* diag;
    uint32_t* first_piv;  // per bottom row: smallest pivot column present (n_top if none)
};

// in-window entries of a pivot row go to the dense diagonal block, not to ent[]
__device__ __forceinline__ bool prep_in_window(const Layout& L, uint32_t row, uint32_t c) {
    return row < L.n_top && c < L.n_top && c < (row / WINDOW + 1) * WINDOW;
}

__global__ void __launch_bounds__(256) prep_count_kernel(PrepArgs a) {
    extern __shared__ uint32_t sh_hist[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_rows = a.L.n_top + a.L.n_bot;
    const uint32_t row = blockIdx.x * 8 + warp;
    uint32_t* hist = sh_hist + warp * a.L.n_bins;
    for (uint32_t b = lane; b < a.L.n_bins; b += 32) hist[b] = 0;
    __syncwarp();
    if (row < n_rows) {
        const uint64_t s = a.row_ptr[row], e = a.row_ptr[row + 1];
        uint32_t fp = a.L.n_top;
        for (uint64_t k = s + lane; k < e; k += 32) {
            const uint32_t c = a.col[k];
            if (c < a.L.n_top && c < fp) fp = c;
            if (prep_in_window(a.L, row, c)) continue;
            const uint32_t t = lay_tile(a.L, c);
            const uint32_t off = c - lay_tile_start(a.L, t);
            atomicAdd(&hist[t * ELIM_WARPS + lay_bucket(off)], 1u);
        }
        __syncwarp();
        for (uint32_t b = lane; b < a.L.n_bins; b += 32) a.seg[(uint64_t)row * a.L.n_bins + b] = hist[b];
        if (row >= a.L.n_top) {
            for (int o = 16; o; o >>= 1) fp = min(fp, __shfl_xor_sync(0xffffffffu, fp, o));
            if (lane == 0) a.first_piv[row - a.L.n_top] = fp;
        }
    }
}

__global__ void __launch_bounds__(256) prep_fill_kernel(PrepArgs a) {
    extern __shared__ uint32_t sh_cur[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_rows = a.L.n_top + a.L.n_bot;
    const uint32_t row = blockIdx.x * 8 + warp;
    if (row >= n_rows) return;
    uint32_t* cur = sh_cur + warp * a.L.n_bins;
    for (uint32_t b = lane; b < a.L.n_bins; b += 32) cur[b] = a.seg[(uint64_t)row * a.L.n_bins + b];
    __syncwarp();
    const uint64_t s = a.row_ptr[row], e = a.row_ptr[row + 1];
    const uint32_t* cf = a.pool + a.coeff_off[a.coeff_id[row]];
    const uint32_t lt = (1u << lane) - 1u;
    for (uint64_t base = s; base < e; base += 32) {
        const uint64_t k = base + lane;
        uint32_t bin = 0xffffffffu, off = 0, val = 0;
        if (k < e) {
            const uint32_t c = a.col[k];
            val = cf[k - s];
            if (a.f.mont_bits) val = fp_mul(val, a.f.rinv_mod_p, a.f);
            if (prep_in_window(a.L, row, c)) {
                if (c != row)  // the unit diagonal is implicit
                    a.diag[((uint64_t)(row / WINDOW) * WINDOW + (row % WINDOW)) * WINDOW + (c % WINDOW)] = val;
            } else {
                const uint32_t t = lay_tile(a.L, c);
                off = c - lay_tile_start(a.L, t);
                bin = t * ELIM_WARPS + lay_bucket(off);
            }
        }
        // stable placement: rank among the lanes of this chunk that share the bin
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (bin != 0xffffffffu) {
            const uint32_t rank = __popc(peers & lt);
            const uint32_t pos = cur[bin] + rank;
            a.ent[pos] = make_uint2(off, val);
        }
        __syncwarp();
        if (bin != 0xffffffffu && (peers >> lane) <= 1u) cur[bin] += __popc(peers);  // highest lane of the group
        __syncwarp();
    }
}

}  // namespace gcu
