// C|D-by-A|B elimination: every row to reduce is eliminated independently
// against the pivot rows (role of the reference's closed kernel/reduce_u16|u32,
// kernel/saxpy, kernel/modular_u16|u32 translation units, CMakeLists.txt:81-90).
//
// One persistent CTA per row in flight, rows pulled from an atomic queue
// (most expensive first). The row's accumulator is a tile of 64-bit lanes in
// shared memory with delayed reduction (field.cuh). The CTA sweeps the column
// tiles left to right:
//   1. zero the tile, scatter the row's own entries of this tile;
//   2. DEFERRED pass: apply every pivot row found in earlier tiles
//      (multiplier list kept per CTA in global memory) restricted to this tile;
//   3. pivot tiles only: walk the tile in WINDOWS of 32 pivot columns. Warp 0
//      solves the window against the dense 32x32 diagonal block in registers
//      (shuffles only; breaks the one-pivot-at-a-time latency chain), publishes
//      the 32 multipliers, and all warps apply the non-zero pivot rows of the
//      window to the rest of the tile, each warp to its own column bucket;
//   4. non-pivot tiles: reduce the lanes mod p and write the row of D'.
// Warps never touch each other's buckets, so no atomics and no intra-window
// barriers are needed; two CTA barriers per window order the window read
// against the previous window's updates.
#pragma once
#include "layout.cuh"

namespace gcu {

struct ElimArgs {
    Layout L;
    Fp f;
    const uint32_t* seg;
    const uint2* ent;
    const uint32_t* diag;
    const uint32_t* first_piv;
    uint32_t* dprime;            // [n_local][n_nonpiv] standard residues
    uint32_t* row_nz;            // [n_local] 1 iff the reduced row is non-zero
    uint32_t* list_k;            // [gridDim][n_top] pivot columns applied so far
    uint32_t* list_m;            // [gridDim][n_top] their multipliers (p - x)
    uint32_t* queue;             // atomic row counter
    uint32_t row_begin, row_step, n_local;  // rows r = row_begin + i * row_step, i < n_local
    const uint64_t* row_ptr;     // block-form CSR row pointers (for the fma_top invariant)
    unsigned long long* stats;   // [0] pivot rows applied, [1] multiply-adds done, [2] fma_top
};

template <bool SMALLP>
__device__ __forceinline__ void elim_rmw(uint64_t* acc, uint2 en, uint32_t m, uint64_t p2) {
    acc[en.x] = fp_mac<SMALLP>(acc[en.x], m, en.y, p2);
}

// Apply up to 32 segments (one per lane: [s,e) with multiplier m) to the tile.
// Loads of four segments are issued before the first read-modify-write so that
// the L2/HBM latency of short segments overlaps.
template <bool SMALLP>
__device__ __forceinline__ void elim_apply_batch(uint64_t* acc, const uint2* __restrict__ ent, uint32_t s,
                                                 uint32_t e, uint32_t m, uint32_t lane, uint64_t p2,
                                                 unsigned long long& fma) {
    uint32_t active = __ballot_sync(0xffffffffu, e > s);
    while (active) {
        int idx[4];
        uint32_t ss[4], ee[4], mm[4];
        uint2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            idx[u] = active ? __ffs(active) - 1 : -1;
            if (active) active &= active - 1;
            const int src = idx[u] < 0 ? 0 : idx[u];
            ss[u] = __shfl_sync(0xffffffffu, s, src);
            ee[u] = __shfl_sync(0xffffffffu, e, src);
            mm[u] = __shfl_sync(0xffffffffu, m, src);
            if (idx[u] < 0) ee[u] = ss[u];
            v[u] = make_uint2(0, 0);
            if (ss[u] + lane < ee[u]) v[u] = __ldg(ent + ss[u] + lane);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (idx[u] < 0) continue;  // warp-uniform
            if (ss[u] + lane < ee[u]) elim_rmw<SMALLP>(acc, v[u], mm[u], p2);
            __syncwarp();
            for (uint32_t q = ss[u] + 32 + lane; q < ee[u]; q += 32) elim_rmw<SMALLP>(acc, __ldg(ent + q), mm[u], p2);
            if (ee[u] - ss[u] > 32) __syncwarp();
            if (lane == 0) fma += ee[u] - ss[u];
        }
    }
}

template <bool SMALLP>
__global__ void __launch_bounds__(ELIM_THREADS, 1) elim_kernel(ElimArgs a) {
    extern __shared__ __align__(16) unsigned char elim_smem[];
    uint64_t* acc = reinterpret_cast<uint64_t*>(elim_smem);
    __shared__ uint32_t s_x[WINDOW];
    __shared__ uint32_t s_nz;
    __shared__ uint32_t s_q;

    const Layout& L = a.L;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t p = a.f.p;
    const uint64_t p2 = a.f.p2;
    uint32_t* lk = a.list_k + (uint64_t)blockIdx.x * L.n_top;
    uint32_t* lm = a.list_m + (uint64_t)blockIdx.x * L.n_top;
    unsigned long long fma = 0, npiv = 0, fma_top = 0;

    for (;;) {
        if (tid == 0) s_q = atomicAdd(a.queue, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        __syncthreads();
        if (q >= a.n_local) break;
        const uint32_t li = a.n_local - 1 - q;                 // later rows have smaller leads: more work
        const uint32_t r = a.row_begin + li * a.row_step;      // index among the rows to reduce
        const uint64_t grow = (uint64_t)(L.n_top + r) * L.n_bins;
        const uint32_t first = a.first_piv[r];
        const uint32_t t_first = first < L.n_top ? first / L.tile_cols : L.n_tiles_piv;
        uint32_t nlist = 0;
        bool nz = false;

        for (uint32_t t = t_first; t < L.n_tiles; ++t) {
            const uint32_t t0 = lay_tile_start(L, t), tw = lay_tile_width(L, t);
            for (uint32_t i = tid; i < tw; i += ELIM_THREADS) acc[i] = 0;
            __syncthreads();
            {   // the row itself
                const uint32_t s = a.seg[grow + t * ELIM_WARPS + warp], e = a.seg[grow + t * ELIM_WARPS + warp + 1];
                for (uint32_t k = s + lane; k < e; k += 32) { const uint2 en = __ldg(a.ent + k); acc[en.x] = en.y; }
                __syncwarp();
            }
            // deferred pass over the pivots found in earlier tiles
            for (uint32_t base = 0; base < nlist; base += 32) {
                const uint32_t i = base + lane;
                uint32_t s = 0, e = 0, m = 0;
                if (i < nlist) {
                    const uint64_t o = (uint64_t)lk[i] * L.n_bins + t * ELIM_WARPS + warp;
                    m = lm[i]; s = a.seg[o]; e = a.seg[o + 1];
                }
                elim_apply_batch<SMALLP>(acc, a.ent, s, e, m, lane, p2, fma);
            }
            __syncthreads();

            if (t < L.n_tiles_piv) {
                const uint32_t nwin = (tw + WINDOW - 1) / WINDOW;
                for (uint32_t g = (t == t_first ? (first - t0) / WINDOW : 0); g < nwin; ++g) {
                    const uint32_t k0 = t0 + g * WINDOW;
                    if (warp == 0) {
                        const uint32_t c = g * WINDOW + lane;
                        uint32_t x = c < tw ? fp_reduce64(acc[c], a.f) : 0u;
                        if (__ballot_sync(0xffffffffu, x != 0)) {
                            const uint32_t* dg = a.diag + (uint64_t)(k0 / WINDOW) * WINDOW * WINDOW + lane;
                            uint32_t d[WINDOW];
#pragma unroll
                            for (int i = 0; i < WINDOW; ++i) d[i] = __ldg(dg + i * WINDOW);
#pragma unroll
                            for (int i = 0; i < WINDOW - 1; ++i) {
                                const uint32_t xi = __shfl_sync(0xffffffffu, x, i);
                                if (xi == 0) continue;  // warp-uniform
                                const uint32_t y = fp_reduce64((uint64_t)x + (uint64_t)(p - xi) * d[i], a.f);
                                if ((int)lane > i) x = y;
                            }
                        }
                        const uint32_t nzm = __ballot_sync(0xffffffffu, x != 0);
                        s_x[lane] = x;
                        if (lane == 0) s_nz = nzm;
                        if (x) {
                            const uint32_t i = nlist + __popc(nzm & lt);
                            lk[i] = k0 + lane; lm[i] = p - x;
                            fma_top += a.row_ptr[k0 + lane + 1] - a.row_ptr[k0 + lane];
                        }
                    }
                    __syncthreads();
                    const uint32_t nzm = s_nz;
                    if (nzm) {
                        uint32_t s = 0, e = 0, m = 0;
                        if ((nzm >> lane) & 1u) {
                            const uint64_t o = (uint64_t)(k0 + lane) * L.n_bins + t * ELIM_WARPS + warp;
                            m = p - s_x[lane]; s = a.seg[o]; e = a.seg[o + 1];
                        }
                        elim_apply_batch<SMALLP>(acc, a.ent, s, e, m, lane, p2, fma);
                        nlist += __popc(nzm);
                    }
                    __syncthreads();
                }
            } else {
                uint32_t* out = a.dprime + (uint64_t)li * L.n_nonpiv + (t0 - L.n_top);
                for (uint32_t i = tid; i < tw; i += ELIM_THREADS) {
                    const uint32_t v = fp_reduce64(acc[i], a.f);
                    out[i] = v;
                    nz |= v != 0;
                }
                __syncthreads();
            }
        }
        const int any = __syncthreads_or(nz ? 1 : 0);
        if (tid == 0) a.row_nz[li] = any ? 1u : 0u;
        npiv += (warp == 0 && lane == 0) ? nlist : 0;
    }
    if (lane == 0) {
        if (fma) atomicAdd(a.stats + 1, fma);
        if (npiv) atomicAdd(a.stats + 0, npiv);
    }
    if (warp == 0 && fma_top) atomicAdd(a.stats + 2, fma_top);
}

}  // namespace gcu
