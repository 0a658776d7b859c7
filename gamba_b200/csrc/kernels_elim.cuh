// C|D-by-A|B elimination: every row to reduce is eliminated independently
// against the pivot rows (role of the reference's closed kernel/reduce_u16|u32,
// kernel/saxpy, kernel/modular_u16|u32 translation units, CMakeLists.txt:81-90).
//
// ONE WARP PER ROW, rows pulled from an atomic queue (most expensive first), as
// many warps resident per SM as shared memory allows: the first version of this
// kernel (one CTA per row, profiles/r01_v1_*) sat at 12 % active warps with 93 %
// L2 hits — the elimination of one row is a latency chain, so throughput comes
// from independent rows in flight, not from more threads per row.
//
// The row's accumulator is a tile of T 64-bit lanes in shared memory with delayed
// reduction (field.cuh). The warp sweeps the column tiles left to right:
//   1. zero the tile, scatter the row's own entries of this tile;
//   2. DEFERRED pass: apply every pivot row found in earlier tiles (multiplier
//      list kept per warp in global memory) restricted to this tile;
//   3. pivot tiles only: walk the tile in WINDOWS of 32 pivot columns. The 32
//      lanes read the window, multiply by the precomputed INVERSE of the 32x32
//      diagonal block (layout.cuh) — no dependent chain: x = a * (I+N)^-1, only
//      rows flagged in dnz are touched — and apply the pivot rows with x != 0 to
//      the rest of the tile, ELIM_DEPTH row segments' loads in flight at a time;
//   4. non-pivot tiles: reduce the lanes mod p and write the row of D'.
// No CTA barriers, no atomics on the accumulator: a warp owns its tile.
#pragma once
#include "layout.cuh"

namespace gcu {

constexpr int ELIM_CTA_WARPS = 4;
constexpr int ELIM_CTA_THREADS = ELIM_CTA_WARPS * 32;

struct ElimArgs {
    Layout L;
    Fp f;
    const uint32_t* seg;
    const uint2* ent;
    const uint32_t* diag;
    const uint32_t* dnz;
    const uint32_t* first_piv;
    uint32_t* dprime;            // [n_local][n_nonpiv] standard residues
    uint32_t* row_nz;            // [n_local] 1 iff the reduced row is non-zero
    uint32_t* list_k;            // [warps in grid][n_top] pivot columns applied so far
    uint32_t* list_m;            // [warps in grid][n_top] their multipliers (p - x)
    uint32_t* queue;             // atomic row counter
    uint32_t row_begin, row_step, n_local;  // rows r = row_begin + i * row_step, i < n_local
    const uint64_t* row_ptr;     // block-form CSR row pointers (for the fma_top invariant)
    unsigned long long* stats;   // [0] pivot rows applied, [1] multiply-adds done, [2] fma_top
};

template <bool SMALLP>
__device__ __forceinline__ void elim_rmw(uint64_t* acc, uint2 en, uint32_t m, uint64_t p2) {
    acc[en.x] = fp_mac<SMALLP>(acc[en.x], m, en.y, p2);
}

// Apply up to 32 segments (lane i holds segment i: entries [s,e) with multiplier m) to the tile.
// The segments are cut into CHUNKS of 32 consecutive entries and the chunk stream of the whole
// batch is software-pipelined: the loads of ELIM_DEPTH chunks are in flight before the first
// read-modify-write. One chunk = all lanes on consecutive entries of ONE pivot row: coalesced,
// and conflict-free because a pivot row has distinct columns. (A row's elimination is a chain of
// L2 round trips; measured variants: 4 segments in flight with unpipelined tails 46 ms,
// lane-per-segment with __match_any_sync conflict checks 320 ms on kat12-15 step 9.)
constexpr int ELIM_DEPTH = 8;

template <bool SMALLP>
__device__ __forceinline__ void elim_apply_batch(uint64_t* acc, const uint2* __restrict__ ent, uint32_t s,
                                                 uint32_t e, uint32_t m, uint32_t lane, uint64_t p2,
                                                 unsigned long long& fma) {
    const uint32_t len = e - s;
    fma += len;
    const uint32_t nch = (len + 31) >> 5;
    uint32_t inc = nch;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if ((int)lane >= o) inc += y; }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    const uint32_t exc = inc - nch;
    for (uint32_t c0 = 0; c0 < total; c0 += ELIM_DEPTH) {
        uint2 v[ELIM_DEPTH];
        uint32_t mm[ELIM_DEPTH];
        bool ok[ELIM_DEPTH];
#pragma unroll
        for (int u = 0; u < ELIM_DEPTH; ++u) {
            const uint32_t c = c0 + u;
            // owner of chunk c: the last lane whose first chunk index is <= c
            const int own = __popc(__ballot_sync(0xffffffffu, exc <= c)) - 1;
            const uint32_t off = (c - __shfl_sync(0xffffffffu, exc, own)) * 32 + lane;
            const uint32_t su = __shfl_sync(0xffffffffu, s, own);
            const uint32_t lu = __shfl_sync(0xffffffffu, len, own);
            mm[u] = __shfl_sync(0xffffffffu, m, own);
            ok[u] = c < total && off < lu;
            v[u] = make_uint2(0, 0);
            if (ok[u]) v[u] = __ldg(ent + su + off);
        }
#pragma unroll
        for (int u = 0; u < ELIM_DEPTH; ++u) {
            if (c0 + u >= total) break;  // warp-uniform
            if (ok[u]) elim_rmw<SMALLP>(acc, v[u], mm[u], p2);
            __syncwarp();
        }
    }
}

template <bool SMALLP>
__global__ void __launch_bounds__(ELIM_CTA_THREADS) elim_kernel(ElimArgs a) {
    extern __shared__ __align__(16) unsigned char elim_smem[];
    const Layout& L = a.L;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* acc = reinterpret_cast<uint64_t*>(elim_smem) + (size_t)warp * L.tile_cols;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t p = a.f.p, nr = L.n_rows;
    const uint64_t p2 = a.f.p2;
    const uint64_t slot = (uint64_t)blockIdx.x * ELIM_CTA_WARPS + warp;
    uint32_t* lk = a.list_k + slot * L.n_top;
    uint32_t* lm = a.list_m + slot * L.n_top;
    unsigned long long fma = 0, npiv = 0, fma_top = 0;

    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(a.queue, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= a.n_local) break;
        const uint32_t li = a.n_local - 1 - q;                 // later rows have smaller leads: more work
        const uint32_t r = a.row_begin + li * a.row_step;      // index among the rows to reduce
        const uint32_t grow = L.n_top + r;
        const uint32_t first = a.first_piv[r];
        const uint32_t t_first = first < L.n_top ? first / L.tile_cols : L.n_tiles_piv;
        uint32_t nlist = 0;
        bool nz = false;

        for (uint32_t t = t_first; t < L.n_tiles; ++t) {
            const uint32_t t0 = lay_tile_start(L, t), tw = lay_tile_width(L, t);
            const uint32_t* seg_s = a.seg + (uint64_t)t * nr;
            const uint32_t* seg_e = seg_s + nr;
            for (uint32_t i = lane; i < tw; i += 32) acc[i] = 0;
            __syncwarp();
            {   // the row itself
                const uint32_t s = seg_s[grow], e = seg_e[grow];
                for (uint32_t k = s + lane; k < e; k += 32) { const uint2 en = __ldg(a.ent + k); acc[en.x] = en.y; }
                __syncwarp();
            }
            // deferred pass over the pivots found in earlier tiles
            for (uint32_t base = 0; base < nlist; base += 32) {
                const uint32_t i = base + lane;
                uint32_t s = 0, e = 0, m = 0;
                if (i < nlist) { const uint32_t k = lk[i]; m = lm[i]; s = seg_s[k]; e = seg_e[k]; }
                elim_apply_batch<SMALLP>(acc, a.ent, s, e, m, lane, p2, fma);
            }

            if (t < L.n_tiles_piv) {
                const uint32_t nwin = (tw + WINDOW - 1) / WINDOW;
                uint32_t g = (t == t_first ? (first - t0) / WINDOW : 0);
                // segment pointers of the window's 32 pivot rows, fetched one window ahead
                uint32_t ps = 0, pe = 0;
                if (g < nwin && t0 + g * WINDOW + lane < L.n_top) { ps = seg_s[t0 + g * WINDOW + lane]; pe = seg_e[t0 + g * WINDOW + lane]; }
                for (; g < nwin; ++g) {
                    const uint32_t k0 = t0 + g * WINDOW;
                    const uint32_t c = g * WINDOW + lane;
                    const uint32_t s = ps, e = pe;
                    if (g + 1 < nwin && k0 + WINDOW + lane < L.n_top) { ps = seg_s[k0 + WINDOW + lane]; pe = seg_e[k0 + WINDOW + lane]; }
                    const uint32_t av = c < tw ? fp_reduce64(acc[c], a.f) : 0u;
                    const uint32_t nza = __ballot_sync(0xffffffffu, av != 0);
                    if (nza == 0) continue;
                    uint32_t x = av;
                    uint32_t need = nza & a.dnz[k0 / WINDOW];
                    if (need) {   // x = a * (I + M): add a_i * M[i][lane] for the flagged rows i
                        const uint32_t* dg = a.diag + (uint64_t)(k0 / WINDOW) * WINDOW * WINDOW + lane;
                        uint64_t xa = av;
                        while (need) {
                            int i0 = __ffs(need) - 1; need &= need - 1;
                            int i1 = need ? __ffs(need) - 1 : -1; if (need) need &= need - 1;
                            const uint32_t m0 = __ldg(dg + i0 * WINDOW);
                            const uint32_t m1 = i1 >= 0 ? __ldg(dg + i1 * WINDOW) : 0u;
                            xa = fp_mac<false>(xa, __shfl_sync(0xffffffffu, av, i0), m0, p2);
                            if (i1 >= 0) xa = fp_mac<false>(xa, __shfl_sync(0xffffffffu, av, i1), m1, p2);
                        }
                        x = fp_reduce64(xa, a.f);
                    }
                    const uint32_t nzm = __ballot_sync(0xffffffffu, x != 0);
                    if (nzm == 0) continue;
                    uint32_t m = 0;
                    if (x) {
                        const uint32_t i = nlist + __popc(nzm & lt);
                        m = p - x;
                        lk[i] = k0 + lane; lm[i] = m;
                        fma_top += a.row_ptr[k0 + lane + 1] - a.row_ptr[k0 + lane];
                    }
                    nlist += __popc(nzm);
                    elim_apply_batch<SMALLP>(acc, a.ent, x ? s : 0u, x ? e : 0u, m, lane, p2, fma);
                }
                __syncwarp();
            } else {
                uint32_t* out = a.dprime + (uint64_t)li * L.n_nonpiv + (t0 - L.n_top);
                for (uint32_t i = lane; i < tw; i += 32) {
                    const uint32_t v = fp_reduce64(acc[i], a.f);
                    out[i] = v;
                    nz |= v != 0;
                }
                __syncwarp();
            }
        }
        const int any = __any_sync(0xffffffffu, nz);
        if (lane == 0) a.row_nz[li] = any ? 1u : 0u;
        npiv += lane == 0 ? nlist : 0;
    }
    for (int o = 16; o; o >>= 1) fma += __shfl_xor_sync(0xffffffffu, fma, o);
    if (lane == 0 && fma) atomicAdd(a.stats + 1, fma);
    if (npiv) atomicAdd(a.stats + 0, npiv);
    if (fma_top) atomicAdd(a.stats + 2, fma_top);
}

}  // namespace gcu
