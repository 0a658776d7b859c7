// Blocked echelon of the dense D' block: up to 32 pivots per pass.
//
// The per-pivot Gauss-Jordan of kernels_echelon.cuh reads and writes every row that
// holds the pivot column once PER PIVOT (measured on cyclic9-15: ~80 MB and 34-47 us
// per pivot, 40 % of the device time of a step). Here a pass works on a PANEL of W <= 32
// consecutive columns [c0, c0+W):
//   P1  CTA 0 loads the panel part of the candidate rows (remaining rows whose lead is
//       inside the panel) into shared memory and runs Gauss-Jordan on it alone: it finds
//       up to W pivots, records for every candidate row r the sequential multipliers
//       M[r][k], keeps a snapshot Qp[k] of each pivot row's panel part as it was when it
//       was used, and writes the reduced panel parts back.
//   -- grid barrier --
//   PC  every CTA replays the pivots on the panel part of its other non-zero rows (the
//       pivots of earlier passes: Jordan part), one warp per row, giving their M[r][k].
//   P2  the trailing parts of the pivot rows "as used" are formed column-parallel:
//       Q_k = D'[p_k] + sum_{k'<k} M[p_k][k'] * Q_k'.
//   -- grid barrier --
//   P3  every CTA updates its rows ONCE for the whole pass: D'[r] += sum_k M[r][k] * Q_k,
//       a (rows x K) by (K x columns) product mod p with the Q tile in shared memory and
//       delayed reduction; recomputes the leads of rows whose panel part vanished and
//       publishes the next panel start with one atomicMin.
//   -- grid barrier --
// A row is therefore read and written once per <= 32 pivots, and there are 3 grid barriers
// per pass instead of one per pivot. The result is the same canonical RREF (pivots in
// ascending column order, every pivot column cleared in all other rows).
#pragma once
#include <cooperative_groups.h>

#include "kernels_echelon.cuh"

namespace gcu {

constexpr int EB_THREADS = 1024;
constexpr int EB_W = 32;                               // panel width = max pivots per pass
constexpr int EB_OWN = 32;                             // owned rows handled per chunk in P3
constexpr int EB_COLS = 1024;                          // columns per Q tile
constexpr uint32_t EB_PENDING = 0xfffffffeu;
constexpr uint32_t EB_CAP = 1024;                      // candidate rows the panel buffer holds at W = 32
constexpr uint32_t EB_MAXC = 8192;                     // most candidate rows a pass can take (W shrinks)
// dynamic shared memory: P1 = list [EB_MAXC] + panel [EB_CAP * 33] + flags [EB_MAXC bytes]; P3 = Q tile + multipliers
constexpr size_t EB_SMEM_P1 = (size_t)EB_MAXC * 4 + (size_t)EB_CAP * (EB_W + 1) * 4 + EB_MAXC;
constexpr size_t EB_SMEM_P3 = ((size_t)EB_W * EB_COLS + (size_t)EB_OWN * EB_W + EB_OWN) * 4;
constexpr size_t EB_SMEM = EB_SMEM_P1 > EB_SMEM_P3 ? EB_SMEM_P1 : EB_SMEM_P3;

struct EbArgs {
    Fp f;
    uint32_t n_rows, n_cols;
    uint32_t* d;
    const uint32_t* row_nz;
    uint32_t* nz_list;              // [n_rows]
    uint32_t* lead;                 // [n_rows] by list index; ECH_INF = pivot / zero, EB_PENDING = to recompute
    uint32_t* mult;                 // [n_rows][EB_W] sequential multipliers of the current pass
    uint32_t* mask;                 // [n_rows] which of them are non-zero
    uint32_t* q;                    // [EB_W][n_cols] pivot rows as used (trailing parts)
    uint32_t* hdr;                  // [0] npk [1] c0 [2] W; [8+k] list index, [40+k] column, [72+k] inverse; [128 + k*32 + j] Qp
    unsigned long long* best;       // [3] rotating (lead << 32 | list index)
    uint32_t* piv_row; uint32_t* piv_col; uint8_t* is_piv_col;
    uint32_t* counts;               // [0] n_nz rows, [1] n_piv
    uint64_t* out_ptr;
};

__device__ __forceinline__ uint32_t eb_block_min(uint32_t v, uint32_t* sh) {
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = sh[threadIdx.x];
        for (int o = 16; o; o >>= 1) w = min(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (threadIdx.x == 0) sh[32] = w;
    }
    __syncthreads();
    return sh[32];
}

__global__ void __launch_bounds__(EB_THREADS, 1) echelon_blocked_kernel(EbArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) uint32_t eb_smem[];
    __shared__ uint32_t sh[33];
    __shared__ uint32_t s_pk[EB_W], s_pc[EB_W], s_inv[EB_W];
    __shared__ uint32_t s_qp[EB_W][EB_W + 1];
    __shared__ uint32_t s_mp[EB_W][EB_W + 1];
    const uint32_t tid = threadIdx.x, nb = gridDim.x, bid = blockIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nr = a.n_rows, nc = a.n_cols, p = a.f.p;
    const uint64_t p2 = a.f.p2;

    // ---- list of the non-zero rows (CTA 0, ballot compaction) ---------------------------------
    if (bid == 0) {
        uint32_t base = 0;
        for (uint32_t r0 = 0; r0 < nr; r0 += EB_THREADS) {
            const uint32_t r = r0 + tid;
            const bool nzr = r < nr && a.row_nz[r] != 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, nzr);
            if (lane == 0) sh[warp] = __popc(bal);
            __syncthreads();
            uint32_t pre = base, tot = 0;
            for (uint32_t w = 0; w < EB_THREADS / 32; ++w) { if (w < warp) pre += sh[w]; tot += sh[w]; }
            if (nzr) a.nz_list[pre + __popc(bal & ((1u << lane) - 1u))] = r;
            base += tot;
            __syncthreads();
        }
        if (tid == 0) a.counts[0] = base;
    }
    for (uint32_t j = bid * EB_THREADS + tid; j < nc; j += nb * EB_THREADS) a.is_piv_col[j] = 0;
    grid.sync();
    const uint32_t nnz_rows = a.counts[0];

    // ---- leads of the owned rows and the first panel start ---------------------------------------
    {
        unsigned long long lb = ~0ull;
        for (uint32_t i = bid; i < nnz_rows; i += nb) {
            const uint32_t* row = a.d + (uint64_t)a.nz_list[i] * nc;
            uint32_t l = ECH_INF;
            for (uint32_t j0 = 0; j0 < nc && l == ECH_INF; j0 += EB_THREADS) {
                const uint32_t j = j0 + tid;
                l = eb_block_min(j < nc && row[j] ? j : ECH_INF, sh);
            }
            if (tid == 0) {
                a.lead[i] = l;
                if (l != ECH_INF) { const unsigned long long k = ((unsigned long long)l << 32) | i; lb = k < lb ? k : lb; }
            }
        }
        if (tid == 0 && lb != ~0ull) atomicMin(a.best + 0, lb);
    }
    grid.sync();

    uint32_t npiv = 0;
    for (uint32_t pass = 0;; ++pass) {
        const unsigned long long key = *(volatile unsigned long long*)(a.best + pass % 3);
        if (key == ~0ull) break;
        const uint32_t c0 = (uint32_t)(key >> 32);

        // ======================= P1: panel factorisation (CTA 0) ==================================
        if (bid == 0) {
            if (tid == 0) a.best[(pass + 2) % 3] = ~0ull;
            // candidates: remaining rows whose lead lies in [c0, c0 + 32), in list order (ordered ballot
            // compaction, so the choice of pivot rows is deterministic)
            uint32_t* cs = eb_smem;                                 // [EB_MAXC] list indices
            uint32_t base = 0;
            for (uint32_t i0 = 0; i0 < nnz_rows; i0 += EB_THREADS) {
                const uint32_t i = i0 + tid;
                const uint32_t l = i < nnz_rows ? a.lead[i] : ECH_INF;
                const bool is = l != ECH_INF && l < c0 + EB_W;
                const uint32_t bal = __ballot_sync(0xffffffffu, is);
                if (lane == 0) sh[warp] = __popc(bal);
                __syncthreads();
                uint32_t pre = base, tot = 0;
                for (uint32_t w = 0; w < EB_THREADS / 32; ++w) { if (w < warp) pre += sh[w]; tot += sh[w]; }
                const uint32_t pos = pre + __popc(bal & ((1u << lane) - 1u));
                if (is && pos < EB_MAXC) cs[pos] = i;
                base += tot;
                __syncthreads();
            }
            const uint32_t ncand = min(base, EB_MAXC);              // the host never launches with more rows
            // panel width: the panel buffer holds EB_CAP rows of 32 columns
            uint32_t W = EB_W;
            while (W > 1 && (uint64_t)ncand * (W + 1) > (uint64_t)EB_CAP * (EB_W + 1)) W >>= 1;
            const uint32_t stride = W + 1;
            uint32_t* S = cs + EB_MAXC;                             // [ncand][W + 1]
            uint8_t* isp = reinterpret_cast<uint8_t*>(S + (size_t)EB_CAP * (EB_W + 1));   // pivot flags [EB_MAXC]
            for (uint32_t x = tid; x < ncand * W; x += EB_THREADS) {
                const uint32_t r = x / W, j = x % W;
                const uint32_t col = c0 + j;
                S[r * stride + j] = col < nc ? a.d[(uint64_t)a.nz_list[cs[r]] * nc + col] : 0u;
            }
            for (uint32_t x = tid; x < ncand; x += EB_THREADS) { isp[x] = 0; a.mask[cs[x]] = 0; }
            __syncthreads();
            uint32_t npk = 0;
            for (uint32_t j = 0; j < W; ++j) {
                uint32_t mine = ECH_INF;
                for (uint32_t r = tid; r < ncand; r += EB_THREADS)
                    if (!isp[r] && S[r * stride + j]) { mine = r; break; }
                const uint32_t pk = eb_block_min(mine, sh);
                if (pk == ECH_INF) continue;                         // CTA-uniform
                const uint32_t inv = fp_inv_fermat(S[pk * stride + j], a.f);   // every thread, no broadcast
                if (tid < EB_W) s_qp[npk][tid] = tid < W ? S[pk * stride + tid] : 0u;
                if (tid == 0) { s_pk[npk] = cs[pk]; s_pc[npk] = c0 + j; s_inv[npk] = inv; isp[pk] = 1; }
                __syncthreads();
                for (uint32_t r = tid; r < ncand; r += EB_THREADS) {
                    if (r == pk) continue;
                    const uint32_t fv = S[r * stride + j];
                    if (fv == 0) continue;
                    const uint32_t m = p - fp_mul(fv, inv, a.f);
                    a.mult[(uint64_t)cs[r] * EB_W + npk] = m;
                    a.mask[cs[r]] |= 1u << npk;
                    for (uint32_t jj = j + 1; jj < W; ++jj) {
                        const uint32_t pv = s_qp[npk][jj];
                        if (pv) S[r * stride + jj] = fp_reduce64((uint64_t)S[r * stride + jj] + (uint64_t)m * pv, a.f);
                    }
                    S[r * stride + j] = 0;
                }
                ++npk;
                __syncthreads();
            }
            // write the reduced panel parts back, set the leads, publish the pass header
            for (uint32_t x = tid; x < ncand * W; x += EB_THREADS) {
                const uint32_t r = x / W, j = x % W;
                if (c0 + j < nc) a.d[(uint64_t)a.nz_list[cs[r]] * nc + c0 + j] = S[r * stride + j];
            }
            for (uint32_t r = tid; r < ncand; r += EB_THREADS) {
                uint32_t l = isp[r] ? ECH_INF : EB_PENDING;
                if (!isp[r]) for (uint32_t j = 0; j < W; ++j) if (S[r * stride + j]) { l = c0 + j; break; }
                a.lead[cs[r]] = l;
            }
            if (tid == 0) { a.hdr[0] = npk; a.hdr[1] = c0; a.hdr[2] = W; }
            if (tid < npk) {
                a.hdr[8 + tid] = s_pk[tid]; a.hdr[40 + tid] = s_pc[tid]; a.hdr[72 + tid] = s_inv[tid];
                a.piv_row[npiv + tid] = a.nz_list[s_pk[tid]]; a.piv_col[npiv + tid] = s_pc[tid]; a.is_piv_col[s_pc[tid]] = 1;
            }
            for (uint32_t x = tid; x < npk * EB_W; x += EB_THREADS) a.hdr[128 + x] = s_qp[x / EB_W][x % EB_W];
        }
        grid.sync();

        // ======================= header, PC: other rows' multipliers ==============================
        const uint32_t npk = a.hdr[0], W = a.hdr[2];
        if (tid < npk) { s_pk[tid] = a.hdr[8 + tid]; s_pc[tid] = a.hdr[40 + tid]; s_inv[tid] = a.hdr[72 + tid]; }
        for (uint32_t x = tid; x < npk * EB_W; x += EB_THREADS) s_qp[x / EB_W][x % EB_W] = a.hdr[128 + x];
        __syncthreads();
        // the pivots' own multipliers (K x K), needed by P2
        for (uint32_t x = tid; x < npk * EB_W; x += EB_THREADS) {
            const uint32_t k = x / EB_W, k2 = x % EB_W;
            s_mp[k][k2] = (k2 < npk && ((a.mask[s_pk[k]] >> k2) & 1u)) ? a.mult[(uint64_t)s_pk[k] * EB_W + k2] : 0u;
        }
        // rows of this CTA that were not candidates but are non-zero (earlier pivots): one warp per row
        for (uint32_t i = bid + warp * nb; i < nnz_rows; i += nb * (EB_THREADS / 32)) {
            const uint32_t l = a.lead[i];
            if (l != ECH_INF) {                                      // candidates (pending) and untouched remaining rows
                if (l != EB_PENDING && lane == 0) a.mask[i] = 0;
                continue;
            }
            bool own_pivot = false;
            for (uint32_t k = 0; k < npk; ++k) own_pivot |= s_pk[k] == i;
            if (own_pivot) continue;                                 // pivots of this pass were candidates
            uint32_t* row = a.d + (uint64_t)a.nz_list[i] * nc;
            const uint32_t col = c0 + lane;
            uint32_t cur = (lane < W && col < nc) ? row[col] : 0u;
            uint32_t msk = 0;
            for (uint32_t k = 0; k < npk; ++k) {
                const uint32_t jk = s_pc[k] - c0;
                const uint32_t fv = __shfl_sync(0xffffffffu, cur, jk);
                if (fv == 0) continue;                               // warp-uniform
                const uint32_t m = p - fp_mul(fv, s_inv[k], a.f);
                if (lane == 0) a.mult[(uint64_t)i * EB_W + k] = m;
                msk |= 1u << k;
                const uint32_t pv = s_qp[k][lane];
                if (lane > jk && pv) cur = fp_reduce64((uint64_t)cur + (uint64_t)m * pv, a.f);
                if (lane == jk) cur = 0;
            }
            if (lane == 0) a.mask[i] = msk;
            if (msk && lane < W && col < nc) row[col] = cur;
        }
        __syncthreads();

        // ======================= P2: pivot rows as used, trailing columns =========================
        const uint32_t ct = c0 + W;                                  // first trailing column
        for (uint32_t j = ct + bid * EB_THREADS + tid; j < nc; j += nb * EB_THREADS) {
            uint32_t qv[EB_W];
#pragma unroll
            for (int k = 0; k < EB_W; ++k) {
                if ((uint32_t)k >= npk) { qv[k] = 0; continue; }
                uint64_t acc = __ldcg(a.d + (uint64_t)a.nz_list[s_pk[k]] * nc + j);
#pragma unroll
                for (int k2 = 0; k2 < k; ++k2) {
                    const uint32_t m = s_mp[k][k2];
                    if (m && qv[k2]) acc = fp_mac<false>(acc, m, qv[k2], p2);
                }
                qv[k] = fp_reduce64(acc, a.f);
                a.q[(uint64_t)k * nc + j] = qv[k];
            }
        }
        grid.sync();

        // ======================= P3: one update of every row for the whole pass ===================
        {
            uint32_t* Qs = eb_smem;                                  // [npk][EB_COLS]
            uint32_t* Ms = eb_smem + EB_W * EB_COLS;                 // [EB_OWN][EB_W]
            uint32_t* s_msk = Ms + EB_OWN * EB_W;                    // [EB_OWN]
            unsigned long long lb = ~0ull;
            for (uint32_t i0 = bid; i0 < nnz_rows; i0 += nb * EB_OWN) {
                __syncthreads();
                if (tid < EB_OWN) { const uint32_t i = i0 + tid * nb; s_msk[tid] = i < nnz_rows ? a.mask[i] : 0u; }
                __syncthreads();
                { const uint32_t k = tid / EB_W, kk = tid % EB_W; const uint32_t i = i0 + k * nb;
                  if (k < EB_OWN) Ms[k * EB_W + kk] = (i < nnz_rows && ((s_msk[k] >> kk) & 1u)) ? a.mult[(uint64_t)i * EB_W + kk] : 0u; }
                uint32_t any = 0;
                for (uint32_t k = 0; k < EB_OWN; ++k) any |= s_msk[k];
                if (any) {
                    for (uint32_t jb = ct; jb < nc; jb += EB_COLS) {
                        __syncthreads();
                        const uint32_t j = jb + tid;
                        for (uint32_t k = 0; k < npk; ++k) Qs[k * EB_COLS + tid] = j < nc ? __ldcg(a.q + (uint64_t)k * nc + j) : 0u;
                        __syncthreads();
                        for (uint32_t k = 0; k < EB_OWN; ++k) {
                            uint32_t msk = s_msk[k];
                            if (msk == 0 || j >= nc) continue;
                            uint32_t* row = a.d + (uint64_t)a.nz_list[i0 + k * nb] * nc;
                            uint64_t acc = __ldcg(row + j);
                            if (a.f.small) {
                                while (msk) { const int kk = __ffs(msk) - 1; msk &= msk - 1; acc += (uint64_t)Ms[k * EB_W + kk] * Qs[kk * EB_COLS + tid]; }
                            } else {
                                while (msk) { const int kk = __ffs(msk) - 1; msk &= msk - 1; acc = fp_mac<false>(acc, Ms[k * EB_W + kk], Qs[kk * EB_COLS + tid], p2); }
                            }
                            row[j] = fp_reduce64(acc, a.f);
                        }
                    }
                }
                // leads of the owned remaining rows: recompute the pending ones, publish the smallest
                __syncthreads();
                for (uint32_t k = 0; k < EB_OWN; ++k) {
                    const uint32_t i = i0 + k * nb;
                    if (i >= nnz_rows) break;
                    uint32_t l = a.lead[i];
                    if (l == ECH_INF) continue;                      // CTA-uniform
                    if (l == EB_PENDING) {
                        const uint32_t* row = a.d + (uint64_t)a.nz_list[i] * nc;
                        l = ECH_INF;
                        for (uint32_t j0 = ct; j0 < nc && l == ECH_INF; j0 += EB_THREADS) {
                            const uint32_t j = j0 + tid;
                            l = eb_block_min(j < nc && __ldcg(row + j) ? j : ECH_INF, sh);
                        }
                        if (tid == 0) a.lead[i] = l;
                    }
                    if (tid == 0 && l != ECH_INF) { const unsigned long long kk = ((unsigned long long)l << 32) | i; lb = kk < lb ? kk : lb; }
                }
            }
            if (tid == 0 && lb != ~0ull) atomicMin(a.best + (pass + 1) % 3, lb);
        }
        npiv += npk;
        grid.sync();
    }
    if (bid == 0 && tid == 0) a.counts[1] = npiv;

    // non-zeros per new row
    for (uint32_t s = bid; s < npiv; s += nb) {
        const uint32_t* row = a.d + (uint64_t)a.piv_row[s] * nc;
        const uint32_t c = a.piv_col[s];
        uint32_t n = 0;
        for (uint32_t j = c + tid; j < nc; j += EB_THREADS) n += (j == c || (!a.is_piv_col[j] && row[j])) ? 1u : 0u;
        for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        __syncthreads();
        if (lane == 0) sh[warp] = n;
        __syncthreads();
        if (tid == 0) { uint32_t t = 0; for (uint32_t w = 0; w < EB_THREADS / 32; ++w) t += sh[w]; a.out_ptr[s + 1] = t; }
    }
    grid.sync();
    if (bid == 0 && tid == 0) {
        uint64_t acc = 0; a.out_ptr[0] = 0;
        for (uint32_t s = 0; s < npiv; ++s) { acc += a.out_ptr[s + 1]; a.out_ptr[s + 1] = acc; }
    }
}

}  // namespace gcu
