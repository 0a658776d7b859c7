// Blocked C-by-A elimination: the sparse triangular solve  M = -C A^-1  for R rows of C|D at a time
// (role of the reference's closed kernel/reduce_u16|u32 translation units, CMakeLists.txt:81-90; used
// together with the tensor-core product of kernels_schur.cuh, which takes the B half).
//
// Why: with one row per CTA (kernels_elim.cuh) a pivot row restricted to one column tile is a piece of
// 4-60 entries, a warp spends ~45 instructions of bookkeeping on each piece and most lanes idle:
// profiles/r01c_elim_*: 1.9 warp instructions per multiply-add, 70 % of the issue slots busy, the same
// pivot rows re-read from L2/HBM by every row (785 GB of DRAM reads per kat14-15 step).
//
// Here a CTA owns R rows (8 for p < 2^16, 4 for p < 2^31) and their accumulator tiles, and the pivot
// rows are applied to all R rows from ONE pass over their entries:
//   * entries are stored tile-major (layout.cuh): the pivot rows of a WINDOW of 32 consecutive pivot
//     columns, restricted to one tile, are one contiguous range, streamed 64 entries per warp
//     instruction (16 B per lane), full lanes whatever the length of the single rows;
//   * every entry carries the slot j of its pivot row inside the window; lane j of the warp holds the
//     R multipliers of pivot 32 w + j, an entry fetches its multipliers with R shuffles;
//   * per entry and row: shuffle, 3 integer instructions (lazy Barrett), one shared-memory RED;
//     rows whose 32 multipliers are all zero are skipped (warp-uniform).
// The multipliers of finished windows live in mt[slot][window][row][32] (global memory, 128-byte
// lines); a bit per window tells whether any of the R rows has a non-zero multiplier there.
// Per tile: zero, scatter the rows' own entries, DEFERRED pass (all earlier windows, dealt to the
// warps), then the tile's own windows one after the other: warp r finds the multipliers of row r
// (x = a (I+N)^-1 with the precomputed inverse diagonal block, as in kernels_elim.cuh), the CTA
// exchanges them through shared memory and all warps share the window's range of entries.
// The non-pivot tiles are not touched: D is scattered straight into the (zeroed) dense D' block.
#pragma once
#include "kernels_elim.cuh"
#include "layout.cuh"

namespace gcu {

#ifndef GAMBA_BLK_WARPS
#define GAMBA_BLK_WARPS 16
#endif
constexpr int BLK_WARPS = GAMBA_BLK_WARPS;
constexpr int BLK_CTAS = 2;                    // CTAs per SM: 32 warps, 64 registers per thread
constexpr int BLK_THREADS = BLK_WARPS * 32;
constexpr int BLK_TMAX = 3200;                 // columns per tile (multiple of WINDOW)
constexpr int BLK_DEPTH = 4;                   // 64-entry chunks a warp keeps in flight
constexpr int BLK_TS = BLK_TMAX + 4;           // accumulator row pitch in columns: + sink lane, 16-byte rows
template <bool SMALLP> struct BlkRows { static constexpr int R = SMALLP ? 8 : 4; };

struct BlkArgs {
    Layout L;
    Fp f;
    uint32_t mu32, negp, norm_every;
    const uint32_t* seg;
    const uint2* ent;
    const uint32_t* diag;
    const uint32_t* dnz;
    const uint32_t* first_piv;
    uint32_t* dprime;            // [n_local][ld_dprime], zeroed by the host: D is scattered into it
    uint32_t* mt;                // [gridDim][n_windows][R][32] multipliers (p - x) of the CTA's current block of rows
    uint32_t* queue;
    uint32_t row_begin, row_step, n_local;
    const uint64_t* row_ptr;
    unsigned long long* stats;   // as ElimArgs
    uint8_t* mx;                 // limb planes of the multipliers for the Schur product
    uint64_t mx_plane;
    uint32_t mx_kp, mx_nl;
    // Monte-Carlo combinations (kernels_mc.cuh): the rows are dense, row li = dense_c[li * dense_ld + pivot column]; their D half
    // is already in dprime. NULL: the rows of C|D themselves
    const uint32_t* dense_c;
    uint64_t dense_ld;
    uint32_t mc_first;           // smallest pivot column any row to reduce holds
};

template <bool SMALLP>
__host__ __device__ constexpr size_t blk_smem_bytes(uint32_t n_windows) {
    return (size_t)BlkRows<SMALLP>::R * BLK_TS * Acc<SMALLP>::WORDS * 4 + (size_t)BlkRows<SMALLP>::R * 32 * 4 +
           2 * WINDOW * WINDOW * 4 + (size_t)((n_windows + 31) / 32 + 1) * 4;
}

// Stream of the entries [s, e) - pivot rows of ONE window restricted to the current tile - through one warp:
// chunks of 64 entries (16 B per lane), chunk index c0, c0 + cstep, ...; BLK_DEPTH chunks are in flight in
// registers. start() only issues loads (it does not need the multipliers: the window walk calls it before
// it works them out); run() applies the entries to the R rows: lane j holds the multipliers m[r] of the
// window's pivot j, an entry fetches the ones of its pivot row with a shuffle per row.
template <bool SMALLP, int R>
struct BlkStream {
    static constexpr int D = BLK_DEPTH;
    static constexpr uint32_t ROWB = BLK_TS * Acc<SMALLP>::WORDS * 4;          // bytes per accumulator row
    uint4 q[D];
    uint32_t pos, cs, e, step;

    __device__ __forceinline__ void start(const uint2* ent, uint32_t s, uint32_t e_, uint32_t c0, uint32_t cstep, uint32_t lane, uint32_t sink) {
        e = e_; step = cstep * 64; cs = s + c0 * 64; pos = cs + lane * 2;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            q[d] = make_uint4(sink, 0u, sink, 0u);
            if (pos < e) q[d] = __ldg(reinterpret_cast<const uint4*>(ent + pos));   // segments hold an even number of entries
            pos += step;
        }
    }
    // DENSE: the multipliers are non-zero (Monte-Carlo combinations): every product is added without a test - a zero only costs a RED
    // of 0 into the column (padding entries: into the sink lane), a test costs a branch around every RED
    template <bool DENSE = false>
    __device__ __forceinline__ unsigned long long run(uint32_t acc_sa, const ElimArgs& ea, const uint2* ent, const uint32_t (&m)[R],
                                                      uint32_t rmask, uint32_t sink) {
        unsigned long long done = 0;
        if (rmask == 0) return 0;
        for (;;) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                if (cs >= e) return done * 64;
                const uint4 cur = q[d];
                q[d] = make_uint4(sink, 0u, sink, 0u);
                if (pos < e) q[d] = __ldg(reinterpret_cast<const uint4*>(ent + pos));
                pos += step; cs += step;
                const uint32_t j0 = (cur.x >> 16) & 31u, j1 = (cur.z >> 16) & 31u;
                const uint32_t a0 = acc_sa + (cur.x & 0xffffu) * (Acc<SMALLP>::WORDS * 4), a1 = acc_sa + (cur.z & 0xffffu) * (Acc<SMALLP>::WORDS * 4);
#pragma unroll
                for (int r = 0; r < R; ++r) {                                  // no branches: the R chains interleave
                    const uint32_t m0 = __shfl_sync(0xffffffffu, m[r], j0), m1 = __shfl_sync(0xffffffffu, m[r], j1);
                    const uint32_t v0 = Acc<SMALLP>::mulred(m0, cur.y, ea), v1 = Acc<SMALLP>::mulred(m1, cur.w, ea);
                    if (DENSE) { Acc<SMALLP>::add(a0 + r * ROWB, 0, v0); Acc<SMALLP>::add(a1 + r * ROWB, 0, v1); }
                    else {
                        Acc<SMALLP>::add_nz(a0 + r * ROWB, v0);                // m v mod p is zero iff m or v is (padding entries: v = 0)
                        Acc<SMALLP>::add_nz(a1 + r * ROWB, v1);
                    }
                }
                done += __popc(rmask);
            }
        }
    }
};

template <bool SMALLP>
__global__ void __launch_bounds__(BLK_THREADS, BLK_CTAS) elim_block_kernel(BlkArgs a) {
    constexpr int R = BlkRows<SMALLP>::R;
    constexpr int W = Acc<SMALLP>::WORDS;
    constexpr uint32_t ROWW = BLK_TS * W;                                  // words per accumulator row
    extern __shared__ __align__(16) unsigned char blk_smem[];
    __shared__ uint32_t s_q, s_first;
    const Layout& L = a.L;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t* acc = reinterpret_cast<uint32_t*>(blk_smem);                 // [R][ROWW]
    uint32_t* xs = acc + R * ROWW;                                         // [R][32] multipliers of the current window
    uint32_t* dgs = xs + R * 32;                                           // [2][32][32] inverse diagonal blocks: this window's, the next one's
    uint32_t* wnz = dgs + 2 * WINDOW * WINDOW;                             // bit w: window w has a non-zero multiplier
    const uint32_t dgs_sa = (uint32_t)__cvta_generic_to_shared(dgs);
    const uint32_t acc_sa = (uint32_t)__cvta_generic_to_shared(acc);
    const uint32_t nww = (L.n_windows + 31) / 32 + 1;
    const uint32_t p = a.f.p, nr = L.n_rows, T = L.tile_cols;
    const uint64_t p2 = a.f.p2;
    const uint32_t n_blocks = (a.n_local + R - 1) / R;
    uint32_t* mt = a.mt + (uint64_t)blockIdx.x * L.n_windows * R * 32;
    ElimArgs ea{};                                                         // the arithmetic helpers of kernels_elim.cuh take these
    ea.f = a.f; ea.mu32 = a.mu32; ea.negp = a.negp;
    unsigned long long fma = 0, npiv = 0, fma_top = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_q = atomicAdd(a.queue, 1u); s_first = 0xffffffffu; }
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= n_blocks) break;
        const uint32_t blk = n_blocks - 1 - q;                             // later rows have smaller leads: more work
        const uint32_t li0 = blk * R;
        if (tid < (uint32_t)R && li0 + tid < a.n_local) atomicMin(&s_first, a.dense_c ? a.mc_first : a.first_piv[a.row_begin + (li0 + tid) * a.row_step]);
        for (uint32_t i = tid; i < nww; i += BLK_THREADS) wnz[i] = 0;
        __syncthreads();
        const uint32_t first = s_first;
        const uint32_t t_first = first < L.n_top ? first / T : L.n_tiles_piv;
        const uint32_t w_first = first / WINDOW;

        for (uint32_t t = t_first; t < L.n_tiles_piv; ++t) {
            const uint32_t t0 = t * T, tw = lay_tile_width(L, t);
            const uint32_t* seg_t = a.seg + (uint64_t)t * nr;
            __syncthreads();
            for (uint32_t i = tid; i < R * ROWW; i += BLK_THREADS) acc[i] = 0;
            __syncthreads();
            // the rows' own entries (C part) of this tile
            if (a.dense_c) {
                for (uint32_t i = tid; i < (uint32_t)R * tw; i += BLK_THREADS) {
                    const uint32_t r = i / tw, c = i - r * tw;
                    if (li0 + r < a.n_local) Acc<SMALLP>::set(acc + r * ROWW, c, __ldcs(a.dense_c + (uint64_t)(li0 + r) * a.dense_ld + t0 + c));
                }
            } else
            for (uint32_t r = warp; r < (uint32_t)R; r += BLK_WARPS) {
                const uint32_t li = li0 + r;
                if (li >= a.n_local) continue;
                const uint32_t grow = L.n_top + a.row_begin + li * a.row_step;
                const uint32_t s = seg_t[grow], e = seg_t[grow + 1];
                for (uint32_t k = s + lane; k < e; k += 32) { const uint2 en = __ldg(a.ent + k); Acc<SMALLP>::set(acc + r * ROWW, en.x & 0xffffu, en.y); }
            }
            __syncthreads();
            // deferred pass: the windows left of this tile, dealt to the warps; lanes are brought back below p
            // before a column can have absorbed norm_every pivot rows
            const uint32_t w_end = t0 / WINDOW;
            const uint32_t w_batch = (a.norm_every - T - 64) / WINDOW;
            for (uint32_t wb = w_first; wb < w_end; wb += w_batch) {
                const uint32_t we = min(w_end, wb + w_batch);
                // multipliers and entry range of a warp's next window are fetched while it applies the current one
                auto next_set = [&](uint32_t w) { while (w < we && !((wnz[w >> 5] >> (w & 31)) & 1u)) w += BLK_WARPS; return w; };
                uint32_t w = next_set(wb + warp);
                uint32_t mn[R], sn = 0, en = 0;
                if (w < we) {
#pragma unroll
                    for (int r = 0; r < R; ++r) mn[r] = __ldcg(mt + ((uint64_t)w * R + r) * 32 + lane);
                    sn = seg_t[w * WINDOW]; en = seg_t[min(w * WINDOW + WINDOW, L.n_top)];
                }
                while (w < we) {
                    uint32_t m[R], rmask = 0;
                    const uint32_t s = sn, e = en;
#pragma unroll
                    for (int r = 0; r < R; ++r) m[r] = mn[r];
                    w = next_set(w + BLK_WARPS);
                    if (w < we) {
#pragma unroll
                        for (int r = 0; r < R; ++r) mn[r] = __ldcg(mt + ((uint64_t)w * R + r) * 32 + lane);
                        sn = seg_t[w * WINDOW]; en = seg_t[min(w * WINDOW + WINDOW, L.n_top)];
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) rmask |= __ballot_sync(0xffffffffu, m[r] != 0) ? 1u << r : 0u;
                    BlkStream<SMALLP, R> st;
                    st.start(a.ent, s, e, 0, 1, lane, T);
                    fma += st.run(acc_sa, ea, a.ent, m, rmask, T);
                }
                __syncthreads();
                for (uint32_t i = tid; i < (uint32_t)R * tw; i += BLK_THREADS) {
                    const uint32_t r = i / tw, c = i - r * tw;
                    Acc<SMALLP>::set(acc + r * ROWW, c, Acc<SMALLP>::get(acc + r * ROWW, c, ea));
                }
                __syncthreads();
            }
            // the tile's own windows
            const uint32_t nwin = (tw + WINDOW - 1) / WINDOW;
            const uint32_t g_first = (t == t_first ? (first - t0) / WINDOW : 0);
            // the inverse diagonal block (4 KB), its row flags and the entry range of a window are fetched one window ahead
            auto stage_diag = [&](uint32_t win, uint32_t buf) {
                cp_async16(dgs_sa + buf * (WINDOW * WINDOW * 4) + tid * 16, a.diag + (uint64_t)win * WINDOW * WINDOW + tid * 4, tid < WINDOW * WINDOW / 4);
                cp_async_commit();
            };
            uint32_t dnz_n = 0, ps_n = 0, pe_n = 0;
            if (g_first < nwin) {
                const uint32_t k0 = t0 + g_first * WINDOW;
                dnz_n = a.dnz[k0 / WINDOW]; ps_n = seg_t[k0]; pe_n = seg_t[min(k0 + WINDOW, L.n_top)];
                stage_diag(k0 / WINDOW, g_first & 1);
            }
            for (uint32_t g = g_first; g < nwin; ++g) {
                const uint32_t k0 = t0 + g * WINDOW, win = k0 / WINDOW;
                const uint32_t c = g * WINDOW + lane;
                const uint32_t dnz_w = dnz_n, ps = ps_n, pe = pe_n;
                cp_async_wait<0>();
                __syncthreads();                                           // this window's diagonal block is in shared memory
                if (g + 1 < nwin) {
                    dnz_n = a.dnz[win + 1]; ps_n = pe; pe_n = seg_t[min(k0 + 2 * WINDOW, L.n_top)];
                    stage_diag(win + 1, (g + 1) & 1);
                }
                BlkStream<SMALLP, R> st;                                   // the window's entries are on their way while the
                st.start(a.ent, ps, pe, warp, BLK_WARPS, lane, T);         // multipliers are worked out
                uint32_t mine = 0, nzm = 0;
                if (warp < (uint32_t)R) {                                  // warp r finds the multipliers of row r
                    const uint32_t av = c < tw ? Acc<SMALLP>::get(acc + warp * ROWW, c, ea) : 0u;
                    const uint32_t nza = __ballot_sync(0xffffffffu, av != 0);
                    uint32_t x = av;
                    uint32_t need = nza & dnz_w;
                    if (need) {   // x = a * (I + M): add a_i * M[i][lane] for the flagged rows i of the inverse block
                        const uint32_t* dg = dgs + (g & 1) * (WINDOW * WINDOW) + lane;
                        uint64_t xa = av;
                        while (need) {
                            const int i0 = __ffs(need) - 1; need &= need - 1;
                            xa = fp_mac<SMALLP>(xa, __shfl_sync(0xffffffffu, av, i0), dg[i0 * WINDOW], p2);
                        }
                        x = fp_reduce64(xa, a.f);
                    }
                    mine = x ? p - x : 0u;
                    nzm = __ballot_sync(0xffffffffu, x != 0);
                    xs[warp * 32 + lane] = mine;
                }
                if (!__syncthreads_or(nzm != 0)) continue;                 // no row of the block has anything in this window
                if (warp < (uint32_t)R) {
                    __stcg(mt + ((uint64_t)win * R + warp) * 32 + lane, mine);
                    const uint32_t li = li0 + warp;
                    if (mine && li < a.n_local) {
                        uint8_t* mp = a.mx + (uint64_t)li * a.mx_kp + k0 + lane;
                        for (uint32_t l = 0; l < a.mx_nl; ++l) mp[l * a.mx_plane] = (uint8_t)(mine >> (8 * l));
                        fma_top += a.row_ptr[k0 + lane + 1] - a.row_ptr[k0 + lane];
                    }
                    if (lane == 0) npiv += __popc(nzm);
                }
                if (tid == 0) wnz[win >> 5] |= 1u << (win & 31);
                uint32_t m[R], rmask = 0;
#pragma unroll
                for (int r = 0; r < R; ++r) { m[r] = xs[r * 32 + lane]; rmask |= __ballot_sync(0xffffffffu, m[r] != 0) ? 1u << r : 0u; }
                // the window's pivot rows, restricted to this tile (their in-window entries are in the diagonal block)
                fma += st.run(acc_sa, ea, a.ent, m, rmask, T);
                // (the barrier at the top of the next window orders these updates before its reads)
            }
        }
        // D part: straight into the zeroed D' block
        if (!a.dense_c)
        for (uint32_t r = warp; r < (uint32_t)R; r += BLK_WARPS) {
            const uint32_t li = li0 + r;
            if (li >= a.n_local) continue;
            const uint32_t grow = L.n_top + a.row_begin + li * a.row_step;
            uint32_t* out = a.dprime + (uint64_t)li * L.ld_dprime;
            for (uint32_t t = L.n_tiles_piv; t < L.n_tiles; ++t) {
                const uint32_t* seg_t = a.seg + (uint64_t)t * nr;
                const uint32_t s = seg_t[grow], e = seg_t[grow + 1], c0 = (t - L.n_tiles_piv) * T;
                for (uint32_t k = s + lane; k < e; k += 32) { const uint2 en = __ldg(a.ent + k); if (en.y) out[c0 + (en.x & 0xffffu)] = en.y; }
            }
        }
    }
    for (int o = 16; o; o >>= 1) { fma += __shfl_xor_sync(0xffffffffu, fma, o); fma_top += __shfl_xor_sync(0xffffffffu, fma_top, o); }
    if (lane == 0 && fma) atomicAdd(a.stats + 1, fma / 32);
    if (lane == 0 && npiv) atomicAdd(a.stats + 0, npiv);
    if (lane == 0 && fma_top) atomicAdd(a.stats + 2, fma_top);
}

}  // namespace gcu
