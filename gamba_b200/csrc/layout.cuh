// Device layout of one staged F4 step (DESIGN.md "Data layout in HBM").
//
// The elimination kernel gives every row to reduce ONE WARP whose accumulator is
// a tile in shared memory (32-bit lanes updated with native shared-memory atomics:
// one lane per column for p < 2^16, a lo/hi pair of lanes per column otherwise). A full-width accumulator does not fit
// (n_cols * 8 B = 0.24 .. 1.6 MB for the named systems), and the
// kernel is latency-bound, so many independent rows must be in flight per SM:
// the column range is cut into TILES of T columns (T * 8 B of shared memory per
// warp) and a warp sweeps the tiles left to right.
//
// Pivot rows (A|B) and rows to reduce (C|D) are therefore stored re-tiled, TILE-MAJOR:
//   ent[]      (tile-relative column | position of the pivot row in its block of 128 rows << 16, value) pairs; all
//              segments of tile 0 in row order, then tile 1, ...: the pivot rows of one window - or of
//              any run of windows - restricted to one tile are ONE contiguous range, which the blocked
//              elimination kernel streams with full lanes whatever the length of the single segments
//              (4-60 entries). Ascending column inside a segment; segments are padded to an even
//              number of entries (16 B) with (tile_cols, 0) - a zero aimed at the tile's sink lane.
//   seg[]      seg[t * n_rows + row] = first entry of (row, tile t); the segment ends where the next
//              element of seg[] starts (one extra element closes the table).
//   diag[]     for every WINDOW of 32 consecutive pivot columns the INVERSE of the
//              dense unit upper-triangular 32x32 block A[w*32+i][w*32+j]; the
//              in-window entries are NOT in ent[]. dnz[w] bit i is set iff row i
//              of the inverse block has an off-diagonal non-zero.
#pragma once
#include <cstdint>

#include "field.cuh"

namespace gcu {

constexpr int WINDOW = 32;                // pivot columns solved together

struct Layout {
    uint32_t n_top, n_bot, n_cols, n_nonpiv, n_rows;
    uint32_t ld_dprime;                   // row pitch of the dense D' block: n_nonpiv rounded up to 4 (16-byte rows)
    uint32_t tile_cols;                   // T, multiple of WINDOW
    uint32_t n_tiles_piv, n_tiles_nonpiv, n_tiles;
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

}  // namespace gcu
