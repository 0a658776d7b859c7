// D' = D + M * B on the tensor cores ("Schur complement" half of the C|D-by-A|B elimination).
//
// The elimination of a row of C|D by the pivot rows A|B is  [0 | D'] = [C | D] + M [A | B]  with the
// multipliers M = -C A^-1. The A half (finding M: a sparse triangular solve, sequential in the pivot
// columns) stays in elim_kernel. The B half is a plain product, and M is DENSE on the named systems
// (tools/step_stats.py: 70-140 k of 160 k pivots hit per row on kat14-15, 60-65 %), while the rows of
// B hold 3-6 % non-zeros: streaming the sparse rows of B once per row to reduce (elim_kernel's
// deferred pass over the non-pivot tiles) moved 785 GB of DRAM traffic per kat14-15 step
// (profiles/r01c) at 0.55e12 multiply-adds/s, bound by shared-memory atomics. As a dense mod-p
// contraction of depth n_top it runs on the tensor cores instead (the north star's "dense mod-p
// contractions" rule, same limb technique as kernels_echelon_mma.cuh):
//
//   Mx[limb][row][k]   u8 limbs of the multipliers, written by the solve (kernels_tri.cuh / elim_kernel) as it finds them
//   Bt[limb][col][k]   u8 limbs of B, transposed (k contiguous), scattered once per step by
//                      schur_bt_kernel from the staged entries of the non-pivot tiles
//   D'[row][col]       u32 residues; holds D when the product starts
//
// The product itself is limb_gemm_tc_kernel (kernels_schur_tc.cuh: tcgen05 + TMA); this file holds the kernels that
// build its operands. (Round 1's mma.sync version of the product was removed: one product kernel, no fallback.)
#pragma once
#include "kernels_echelon_mma.cuh"
#include "kernels_elim.cuh"
#include "layout.cuh"

namespace gcu {

constexpr int SC_BK = 64;                  // bytes of k per stage and row


struct BtArgs {
    Layout L;
    const uint32_t* seg;
    const uint2* ent;
    uint8_t* bt;                // [nl][np][kp]
    uint64_t bt_plane;
    uint32_t kp, nl;
    // panel mode: the part of A right of a pivot row's own tile goes to per-tile operands as well:
    // at + at_off[t] is [nl][tile_cols][kp_t] with kp_t = t * tile_cols (only pivots of earlier tiles reach tile t)
    uint8_t* at;                // NULL: B only
    const uint64_t* at_off;     // [n_tiles_piv] byte offsets
    // blocked triangular solve (kernels_tri.cuh): the part of A inside a pivot row's own tile, transposed:
    // adt[l][g][k'] = A[tile start + k'][g] for the columns g right of the row's window, [nl][kt][tile_cols]
    uint8_t* adt;               // NULL: not needed
    uint64_t adt_plane;
};

// One warp per pivot row k: its entries in the non-pivot tiles become column k of Bt. Consecutive
// warps write adjacent bytes of the same sectors, which L2 merges before they reach HBM.
__global__ void __launch_bounds__(256) schur_bt_kernel(BtArgs a) {
    const Layout& L = a.L;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t k = blockIdx.x * 8 + warp;
    if (k >= L.n_top) return;
    const uint32_t nr = L.n_rows;
    if (a.adt) {
        const uint32_t t = k / L.tile_cols, T = L.tile_cols;
        const uint32_t s = a.seg[(uint64_t)t * nr + k], e = a.seg[(uint64_t)t * nr + k + 1];
        uint8_t* base = a.adt + (uint64_t)t * T * T + (k - t * T);
        for (uint32_t i = s + lane; i < e; i += 32) {
            const uint2 en = __ldg(a.ent + i);
            if (en.y == 0) continue;                       // padding
            uint8_t* dst = base + (uint64_t)(en.x & 0xffffu) * T;
            for (uint32_t l = 0; l < a.nl; ++l) dst[l * a.adt_plane] = (uint8_t)(en.y >> (8 * l));
        }
    }
    if (a.at) {
        for (uint32_t t = k / L.tile_cols + 1; t < L.n_tiles_piv; ++t) {
            const uint32_t s = a.seg[(uint64_t)t * nr + k], e = a.seg[(uint64_t)t * nr + k + 1];
            const uint64_t kpt = (uint64_t)t * L.tile_cols, plane = kpt * L.tile_cols;
            uint8_t* base = a.at + a.at_off[t] + k;
            for (uint32_t i = s + lane; i < e; i += 32) {
                const uint2 en = __ldg(a.ent + i);
                if (en.y == 0) continue;                   // padding
                uint8_t* dst = base + (uint64_t)(en.x & 0xffffu) * kpt;
                for (uint32_t l = 0; l < a.nl; ++l) dst[l * plane] = (uint8_t)(en.y >> (8 * l));
            }
        }
    }
    for (uint32_t t = L.n_tiles_piv; t < L.n_tiles; ++t) {
        const uint32_t s = a.seg[(uint64_t)t * nr + k], e = a.seg[(uint64_t)t * nr + k + 1];
        const uint32_t c0 = (t - L.n_tiles_piv) * L.tile_cols;
        for (uint32_t i = s + lane; i < e; i += 32) {
            const uint2 en = __ldg(a.ent + i);
            if (en.y == 0) continue;                       // padding
            uint8_t* dst = a.bt + (uint64_t)(c0 + (en.x & 0xffffu)) * a.kp + k;
            for (uint32_t l = 0; l < a.nl; ++l) dst[l * a.bt_plane] = (uint8_t)(en.y >> (8 * l));
        }
    }
}

// ---- Monte-Carlo combinations of the rows of D' before the echelon (README.md:5 of the reference: "Monte-Carlo F4") ----
// K random combinations  Dc = S D'  (S: K x n_rows, uniform limb bytes) span the row space of D' except with probability
// ~p^-(K - rank); the echelon then runs on K rows instead of n_rows (85-90 % of which would reduce to zero). The product
// is one more use of the limb-split tensor-core kernel: S limbs as the left operand, D' limbs transposed as the right one.
struct McArgs {
    uint32_t n_rows, n_cols, ld;    // D' : n_rows x n_cols, pitch ld
    const uint32_t* d;
    const uint32_t* row_nz;         // rows flagged zero are skipped (padding rows of the all-gather buffer hold stale data)
    uint8_t* dt;                    // [nl][np][rp]  limbs of D' transposed (row index contiguous)
    uint8_t* s;                     // [nl][kp][rp]  limbs of the random multipliers
    uint64_t dt_plane, s_plane;
    uint32_t rp, k, nl;
    uint64_t seed;
};

// 32 columns x 128 rows of D' per CTA (256 threads): coalesced reads along the columns, 4 row-bytes per store
__global__ void __launch_bounds__(256) mc_dt_kernel(McArgs a) {
    __shared__ uint32_t tile[128][33];
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 128;
    for (uint32_t i = ty; i < 128; i += 8) {
        const uint32_t r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < a.n_rows && c < a.n_cols && a.row_nz[r]) ? a.d[(uint64_t)r * a.ld + c] : 0u;
    }
    __syncthreads();
    for (uint32_t w = threadIdx.x; w < 32 * 32; w += 256) {
        const uint32_t cl = w >> 5, word = w & 31;
        if (c0 + cl >= a.n_cols) continue;
        const uint32_t v0 = tile[word * 4][cl], v1 = tile[word * 4 + 1][cl], v2 = tile[word * 4 + 2][cl], v3 = tile[word * 4 + 3][cl];
        for (uint32_t l = 0; l < a.nl; ++l) {
            const uint32_t sh = 8 * l;
            const uint32_t packed = ((v0 >> sh) & 0xffu) | (((v1 >> sh) & 0xffu) << 8) | (((v2 >> sh) & 0xffu) << 16) | (((v3 >> sh) & 0xffu) << 24);
            *reinterpret_cast<uint32_t*>(a.dt + l * a.dt_plane + (uint64_t)(c0 + cl) * a.rp + r0 + word * 4) = packed;
        }
    }
}

// random limb bytes of the K x n_rows multipliers (columns >= n_rows of the padded planes stay zero)
__global__ void __launch_bounds__(256) mc_s_kernel(McArgs a) {
    const uint64_t words_per_row = a.rp / 4;
    const uint64_t idx = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (uint64_t)a.k * words_per_row) return;
    const uint32_t q = (uint32_t)(idx / words_per_row), r4 = (uint32_t)(idx % words_per_row) * 4;
    for (uint32_t l = 0; l < a.nl; ++l) {
        uint64_t z = a.seed + 0x9e3779b97f4a7c15ull * ((((uint64_t)q << 32) | r4) * 4 + l) + 0x632be59bd9b4e019ull;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z ^= z >> 31;
        uint32_t w = (uint32_t)z;
        if (l == a.nl - 1) w &= 0x7f7f7f7fu;           // top limb below 128: the multiplier stays below 2^(8 nl - 1) (any value works mod p)
        uint32_t keep = 0;
        for (uint32_t b = 0; b < 4; ++b) if (r4 + b < a.n_rows) keep |= 0xffu << (8 * b);
        *reinterpret_cast<uint32_t*>(a.s + l * a.s_plane + (uint64_t)q * a.rp + r4) = w & keep;
    }
}


}  // namespace gcu
