// Device layout of one staged F4 step (DESIGN.md "Data layout in HBM").
//
// The elimination kernel gives every row to reduce one CTA whose accumulator
// lives in shared memory. A full-width 64-bit accumulator does not fit
// (n_cols * 8 B = 0.24 .. 1.6 MB for the named systems), so the
// column range is cut into TILES of T columns; a CTA sweeps the tiles left to
// right. Inside a tile the ELIM_WARPS warps own disjoint column BUCKETS
// (16-column blocks dealt round-robin to the warps), which makes every
// read-modify-write of the accumulator conflict-free without atomics.
//
// Pivot rows (A|B) and rows to reduce (C|D) are therefore stored re-bucketed:
//   ent[]      (tile-relative column, value) pairs, grouped per row by
//              bin = tile * ELIM_WARPS + bucket, ascending column inside a bin
//   seg[]      seg[row * n_bins + bin] = first entry of that bin; bins of a row
//              and rows are contiguous, so the end is the next seg value
//   diag[]     for every window of 32 consecutive pivot columns the dense
//              32x32 block A[w*32+i][w*32+j]; those entries are NOT in ent[]
//              (the window is solved in registers, kernels_elim.cuh)
#pragma once
#include <cstdint>

#include "field.cuh"

namespace gcu {

constexpr int ELIM_WARPS = 8;             // warps per elimination CTA == buckets per tile
constexpr int ELIM_THREADS = ELIM_WARPS * 32;
constexpr int BUCKET_COLS = 16;           // 16 x 8 B = one 128 B shared-memory row per bucket block
constexpr int WINDOW = 32;                // pivot columns solved together

struct Layout {
    uint32_t n_top, n_bot, n_cols, n_nonpiv;
    uint32_t tile_cols;                   // T, multiple of BUCKET_COLS * ELIM_WARPS and of WINDOW
    uint32_t n_tiles_piv, n_tiles_nonpiv, n_tiles, n_bins;
    uint32_t n_windows;                   // ceil(n_top / 32)
};

__host__ __device__ __forceinline__ uint32_t lay_tile(const Layout& L, uint32_t c) {
    return c < L.n_top ? c / L.tile_cols : L.n_tiles_piv + (c - L.n_top) / L.tile_cols;
}
__host__ __device__ __forceinline__ uint32_t lay_tile_start(const Layout& L, uint32_t t) {
    return t < L.n_tiles_piv ? t * L.tile_cols : L.n_top + (t - L.n_tiles_piv) * L.tile_cols;
}
__host__ __device__ __forceinline__ uint32_t lay_tile_width(const Layout& L, uint32_t t) {
    uint32_t s = lay_tile_start(L, t);
    uint32_t end = t < L.n_tiles_piv ? L.n_top : L.n_cols;
    uint32_t w = end - s;
    return w < L.tile_cols ? w : L.tile_cols;
}
__host__ __device__ __forceinline__ uint32_t lay_bucket(uint32_t off_in_tile) {
    return (off_in_tile / BUCKET_COLS) % ELIM_WARPS;
}

}  // namespace gcu
