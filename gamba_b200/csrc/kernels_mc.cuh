// Monte-Carlo random linear combinations of the rows to reduce (README.md:5 of the reference: "Monte-Carlo F4";
// the seed of linalg_v3{field, seed}, source/f4.hpp:66): instead of the n_bot rows of C|D the engine eliminates
//   Cc|Dc = S (C|D),   S: K x n_bot uniform over F_p,   K = predicted rank of D' + margin,
// K rows instead of n_bot (katsura / eco: 85-92 % of the rows reduce to zero). The reduced combinations span the row
// space of D' except with probability ~p^-(K - rank); they are accepted only with K - rank >= 32 (run_reduce), so the
// canonical RREF the echelon returns is that of the exact computation.
//
// The combinations are dense rows and are formed ONCE, in front of whichever solve the step uses:
//   mc_gen_kernel    S[q][row], uniform in [0, p)
//   mc_rows_kernel   one CTA per (R combinations, column tile): the tile's entries of ALL rows to reduce - one
//                    contiguous range of ent[] (layout.cuh) - streamed window by window (32 rows) through the warps
//                    with the machinery of the blocked elimination (BlkStream: 64 entries per warp instruction,
//                    lane j holds the R multipliers of row 32 w + j, one shared-memory RED per entry and combination);
//                    the accumulator tile leaves as dense u32 rows: pivot tiles into Cc, non-pivot tiles into D'.
// Work: K x nnz(C|D) multiply-adds, 5-10 % of what the elimination of the K rows costs.
#pragma once
#include "kernels_elim_block.cuh"

namespace gcu {

// multiplier of row r in combination q: uniform in [0, p) (splitmix64 finaliser)
__device__ __forceinline__ uint32_t mc_coeff(uint64_t seed, uint32_t q, uint32_t r, uint32_t p) {
    uint64_t z = seed + 0x9e3779b97f4a7c15ull * (((uint64_t)q << 32) | r) + 0x632be59bd9b4e019ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    return (uint32_t)(z % p);
}

struct McRowsArgs {
    Layout L;
    Fp f;
    uint32_t mu32, negp, norm_every;
    const uint32_t* seg;
    const uint2* ent;
    uint32_t* s;                 // [n_q_pad][nbp] multipliers; rows q >= n_q and columns >= n_bot are zero
    uint32_t nbp;                // n_bot rounded up to a multiple of 32
    uint32_t n_q, n_q_pad;       // combinations of this rank, rounded up to a multiple of R
    uint32_t q_global0;          // number of the first one among all combinations drawn for this step
    uint64_t seed;
    uint32_t* dst_c;             // pivot tiles:     dst_c[q * ldc + column], every column < ldc written
    uint64_t ldc;                //                  n_tiles_piv * tile_cols
    uint32_t* dst_d;             // non-pivot tiles: dst_d[q * ldd + column - n_top]
    uint32_t ldd;
};

__global__ void __launch_bounds__(256) mc_gen_kernel(McRowsArgs a) {
    const uint64_t idx = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (uint64_t)a.n_q_pad * a.nbp) return;
    const uint32_t q = (uint32_t)(idx / a.nbp), row = (uint32_t)(idx % a.nbp);
    a.s[idx] = (q < a.n_q && row < a.L.n_bot) ? mc_coeff(a.seed, a.q_global0 + q, row, a.f.p) : 0u;
}

template <bool SMALLP>
__global__ void __launch_bounds__(BLK_THREADS, BLK_CTAS) mc_rows_kernel(McRowsArgs a) {
    constexpr int R = BlkRows<SMALLP>::R;
    constexpr int W = Acc<SMALLP>::WORDS;
    constexpr uint32_t ROWW = BLK_TS * W;
    extern __shared__ __align__(16) unsigned char mcr_smem[];
    uint32_t* acc = reinterpret_cast<uint32_t*>(mcr_smem);                 // [R][ROWW]
    const uint32_t acc_sa = (uint32_t)__cvta_generic_to_shared(acc);
    const Layout& L = a.L;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t q0 = blockIdx.x * R, t = blockIdx.y;
    const uint32_t T = L.tile_cols, nr = L.n_rows, tw = lay_tile_width(L, t), t0 = lay_tile_start(L, t);
    const uint32_t* seg_t = a.seg + (uint64_t)t * nr;
    ElimArgs ea{};
    ea.f = a.f; ea.mu32 = a.mu32; ea.negp = a.negp;
    for (uint32_t i = tid; i < R * ROWW; i += BLK_THREADS) acc[i] = 0;
    __syncthreads();
    const uint32_t n_win = (L.n_bot + WINDOW - 1) / WINDOW;
    const uint32_t w_batch = (a.norm_every - 64) / WINDOW;                 // a column absorbs at most one entry per row and combination
    for (uint32_t wb = 0; wb < n_win; wb += w_batch) {
        const uint32_t we = min(n_win, wb + w_batch);
        for (uint32_t w = wb + warp; w < we; w += BLK_WARPS) {
            uint32_t m[R], rmask = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) m[r] = __ldg(a.s + (uint64_t)(q0 + r) * a.nbp + w * WINDOW + lane);
            const uint32_t s = seg_t[L.n_top + w * WINDOW], e = seg_t[min(L.n_top + w * WINDOW + WINDOW, nr)];
#pragma unroll
            for (int r = 0; r < R; ++r) rmask |= __ballot_sync(0xffffffffu, m[r] != 0) ? 1u << r : 0u;
            BlkStream<SMALLP, R> st;
            st.start(a.ent, s, e, 0, 1, lane, T);
            st.template run<true>(acc_sa, ea, a.ent, m, rmask, T);
        }
        __syncthreads();
        for (uint32_t i = tid; i < (uint32_t)R * tw; i += BLK_THREADS) {
            const uint32_t r = i / tw, c = i - r * tw;
            Acc<SMALLP>::set(acc + r * ROWW, c, Acc<SMALLP>::get(acc + r * ROWW, c, ea));
        }
        __syncthreads();
    }
    // dense rows out: every column of the tile (pivot tiles: up to the padded width, zeros beyond the tile's own columns)
    const bool piv = t < L.n_tiles_piv;
    const uint32_t wout = piv ? T : tw;
    for (uint32_t i = tid; i < (uint32_t)R * wout; i += BLK_THREADS) {
        const uint32_t r = i / wout, c = i - r * wout, q = q0 + r;
        if (q >= a.n_q) break;
        const uint32_t v = c < tw ? Acc<SMALLP>::get(acc + r * ROWW, c, ea) : 0u;
        if (piv) a.dst_c[(uint64_t)q * a.ldc + t0 + c] = v;
        else a.dst_d[(uint64_t)q * a.ldd + (t0 - L.n_top) + c] = v;
    }
}

}  // namespace gcu
