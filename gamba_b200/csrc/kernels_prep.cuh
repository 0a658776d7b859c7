// Staging kernels: turn the block-form CSR handed over at the ABI (column lists +
// shared coefficient arrays) into the tiled device layout of
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
    uint32_t* seg;        // [n_tiles * n_rows + 1] tile-major: counts, then (after the exclusive scan) segment starts;
                          // (row, tile) ends where the next element starts (layout.cuh)
    uint2* ent;
    uint32_t* diag;       // [n_windows][32][32] strictly upper part of the diagonal blocks of A (raw: kernels_tri.cuh reads it)
    uint32_t* dinv;       // [n_windows][32][32] strictly upper part of their inverses (what the elimination kernels read)
    uint32_t* dnz;        // [n_windows]
    uint32_t* first_piv;  // per bottom row: smallest pivot column present (n_top if none)
    uint32_t* first_min;  // min of first_piv over all rows to reduce (preset to 0xffffffff)
    const uint64_t* row_src;  // optional: where row r's entries start in col[] (default row_ptr[r])
    const uint32_t* col_map;  // optional: col[] holds handles, block-form column = col_map[handle]
};

// in-window entries of a pivot row go to the dense diagonal block, not to ent[]
__device__ __forceinline__ bool prep_in_window(const Layout& L, uint32_t row, uint32_t c) {
    return row < L.n_top && c < L.n_top && c < (row / WINDOW + 1) * WINDOW;
}

__global__ void __launch_bounds__(256) prep_count_kernel(PrepArgs a) {
    extern __shared__ uint32_t sh_hist[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * 8 + warp;
    const uint32_t nt = a.L.n_tiles;
    uint32_t* hist = sh_hist + warp * nt;
    for (uint32_t b = lane; b < nt; b += 32) hist[b] = 0;
    __syncwarp();
    if (row < a.L.n_rows) {
        const uint64_t s = a.row_src ? a.row_src[row] : a.row_ptr[row], e = s + (a.row_ptr[row + 1] - a.row_ptr[row]);
        uint32_t fp = a.L.n_top;
        for (uint64_t k = s + lane; k < e; k += 32) {
            const uint32_t c = a.col_map ? a.col_map[a.col[k]] : a.col[k];
            if (c < a.L.n_top && c < fp) fp = c;
            if (prep_in_window(a.L, row, c)) continue;
            atomicAdd(&hist[lay_tile(a.L, c)], 1u);
        }
        __syncwarp();
        for (uint32_t b = lane; b < nt; b += 32) a.seg[(uint64_t)b * a.L.n_rows + row] = (hist[b] + 1u) & ~1u;  // padded to 16 B
        if (row >= a.L.n_top) {
            for (int o = 16; o; o >>= 1) fp = min(fp, __shfl_xor_sync(0xffffffffu, fp, o));
            if (lane == 0) { a.first_piv[row - a.L.n_top] = fp; atomicMin(a.first_min, fp); }
        }
    }
}

__global__ void __launch_bounds__(256) prep_fill_kernel(PrepArgs a) {
    extern __shared__ uint32_t sh_cur[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * 8 + warp;
    const uint32_t nt = a.L.n_tiles, nr = a.L.n_rows;
    if (row >= nr) return;
    uint32_t* cur = sh_cur + warp * nt;
    for (uint32_t b = lane; b < nt; b += 32) cur[b] = a.seg[(uint64_t)b * nr + row];
    __syncwarp();
    // position of the row in its block of 128 pivot rows / of 128 rows to reduce (low 5 bits: in its window of 32)
    const uint32_t jbits = ((row < a.L.n_top ? row : row - a.L.n_top) % 128u) << 16;
    const uint64_t s = a.row_src ? a.row_src[row] : a.row_ptr[row], e = s + (a.row_ptr[row + 1] - a.row_ptr[row]);
    const uint32_t* cf = a.pool + a.coeff_off[a.coeff_id[row]];
    const uint32_t lt = (1u << lane) - 1u;
    for (uint64_t base = s; base < e; base += 32) {
        const uint64_t k = base + lane;
        uint32_t bin = 0xffffffffu, off = 0, val = 0;
        if (k < e) {
            const uint32_t c = a.col_map ? a.col_map[a.col[k]] : a.col[k];
            val = cf[k - s];
            if (a.f.mont_bits) val = fp_mul(val, a.f.rinv_mod_p, a.f);
            if (prep_in_window(a.L, row, c)) {
                if (c != row)  // the unit diagonal is implicit
                    a.diag[((uint64_t)(row / WINDOW) * WINDOW + (row % WINDOW)) * WINDOW + (c % WINDOW)] = val;
            } else {
                bin = lay_tile(a.L, c);
                off = c - lay_tile_start(a.L, bin);
            }
        }
        // stable placement: rank among the lanes of this chunk that share the tile
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (bin != 0xffffffffu) {
            const uint32_t pos = cur[bin] + __popc(peers & lt);
            a.ent[pos] = make_uint2(off | jbits, val);
        }
        __syncwarp();
        if (bin != 0xffffffffu && (peers >> lane) <= 1u) cur[bin] += __popc(peers);  // highest lane of the group
        __syncwarp();
    }
    // pad every (row, tile) segment to an even number of entries (16 B) with a zero-valued entry aimed at
    // the sink lane (index tile_cols) of the accumulator tile
    for (uint32_t b = lane; b < nt; b += 32)
        if (cur[b] & 1u) a.ent[cur[b]] = make_uint2(a.L.tile_cols | jbits, 0u);
}

// Number of (padded) entries of the pivot rows inside the pivot tiles = nnz of A without the diagonal blocks:
// per tile the pivot rows' entries are the contiguous range [seg[t*nr], seg[t*nr + n_top]). One warp.
__global__ void prep_nnz_a_kernel(PrepArgs a, unsigned long long* out) {
    unsigned long long s = 0;
    for (uint32_t t = threadIdx.x; t < a.L.n_tiles_piv; t += 32)
        s += a.seg[(uint64_t)t * a.L.n_rows + a.L.n_top] - a.seg[(uint64_t)t * a.L.n_rows];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *out = s;
}

// One warp per window: from the strictly-upper block N (A_ww = I + N, diagonal implicit) the
// strictly-upper part M of (I + N)^-1 = I + M, lane = column:  M[i] = -N[i] - sum_{j>i} N[i][j] * M[j].
__global__ void __launch_bounds__(256) prep_diag_inverse_kernel(PrepArgs a) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * 8 + warp;
    if (w >= a.L.n_windows) return;
    const uint32_t* blk = a.diag + (uint64_t)w * WINDOW * WINDOW;
    uint32_t* out = a.dinv + (uint64_t)w * WINDOW * WINDOW;
    uint32_t n[WINDOW], inv[WINDOW];
#pragma unroll
    for (int i = 0; i < WINDOW; ++i) { n[i] = blk[i * WINDOW + lane]; inv[i] = 0; }
    uint32_t mask = 0;
#pragma unroll
    for (int i = WINDOW - 2; i >= 0; --i) {
        if (__ballot_sync(0xffffffffu, n[i] != 0) == 0) continue;  // empty row: inverse row is e_i
        uint64_t acc = n[i] ? (uint64_t)(a.f.p - n[i]) : 0ull;
#pragma unroll
        for (int j = i + 1; j < WINDOW; ++j) {
            const uint32_t nij = __shfl_sync(0xffffffffu, n[i], j);
            if (nij == 0) continue;  // warp-uniform
            if (inv[j]) acc = fp_reduce64(acc + (uint64_t)(a.f.p - nij) * inv[j], a.f);
        }
        inv[i] = (int)lane > i ? fp_reduce64(acc, a.f) : 0u;
        if (__ballot_sync(0xffffffffu, inv[i] != 0)) mask |= 1u << i;
    }
#pragma unroll
    for (int i = 0; i < WINDOW; ++i) out[i * WINDOW + lane] = inv[i];
    if (lane == 0) a.dnz[w] = mask;
}

}  // namespace gcu
