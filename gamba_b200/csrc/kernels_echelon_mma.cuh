// Echelon of the dense D' block on the tensor cores (role of the reference's closed
// kernel/echelon_u16|u32 translation units, CMakeLists.txt:81-90).
//
// Blocked Gauss-Jordan, ONE cooperative kernel, two grid barriers per ROUND of up to 32 pivots
// (the per-pivot kernel of kernels_echelon.cuh pays one barrier and one full pass over D' per pivot:
// 36 us per pivot on cyclic9-15, bound by re-reading D'):
//
//   phase 1 (every CTA redundantly, in shared memory, no barrier): take EM_B = 32 remaining rows
//           whose leads fall into a WINDOW of EM_W columns starting at the smallest lead, load the
//           32 x EM_W window, run Gauss-Jordan on it fraction-free (row_i = d*row_i - m*row_piv:
//           no modular inverse on the critical path) while tracking the row operations in an
//           appended 32 x 32 block T; one batch of <= 32 inverses at the end makes T monic.
//           Result: pivot columns c_1..c_k (k <= 32) and T with  NewPivotRows = T * PanelRows.
//   phase 2 (grid): every live row r gets its multipliers x_j = D'[r][c_j] and its row of the
//           update operand  X'[r] = -(x * T) mod p  (new pivot rows: X' = T[j], old content dropped);
//           X' and the 32 panel rows are split into 8-bit limbs (2 limbs for p < 2^16, 4 for p < 2^31).
//   phase 3 (grid): trailing update  D'[live rows] += X' * PanelRows  for all columns right of the
//           window start: a mod-p contraction of depth 32 on the tensor cores. int8 IMMA
//           (mma.sync.m16n8k32.u8.u8.s32) on the limb planes: products of limbs i and j accumulate
//           exactly in s32 (32 * 255^2 * 4 < 2^24) into the accumulator of weight 2^(8(i+j)); the
//           epilogue recombines the 2*NL-1 accumulators with two Barrett reductions, adds the old
//           entry, stores, and records the new lead of the row.
//
// Why IMMA and not FP64 DMMA (the north star leaves the choice to measurement): tools/micro/mma_bench.cu
// on B200: DMMA m8n8k4 18.5 TMAC/s, IMMA m16n8k32 570 TMAC/s. A mod-p MAC costs 1 DMMA MAC (p < 2^16)
// or 2 (p < 2^31, 16-bit split with k-blocks of 64) against 4 or 16 IMMA MACs: IMMA is 7.7x / 3.9x
// faster. (tcgen05.mma kind::i8 would double that again; with depth-32 updates the kernel is bound by
// the traffic on D', not by the tensor pipe, see DESIGN.md.)
//
// Any pivoting order gives the same result: once every pivot column is cleared from all other rows
// and rows are sorted by pivot, the matrix is THE reduced row echelon form of the row space.
#pragma once
#include <cooperative_groups.h>

#include "kernels_echelon.cuh"
#include "layout.cuh"

namespace gcu {

constexpr int EM_THREADS = 512;
constexpr int EM_WARPS = EM_THREADS / 32;
constexpr int EM_B = 32;                    // panel rows = contraction depth of one trailing update
#ifndef GAMBA_EM_W
#define GAMBA_EM_W 256          // measured on the cyclic9-15 batch (tools/emw_variants.sh): 512: 29.5 ms, 256: 26.9 ms, 128: 29.5 ms per pass
#endif
constexpr int EM_W = GAMBA_EM_W;                   // window columns of the panel factorisation
constexpr int EM_WP = EM_W + EM_B;          // + the transformation block
constexpr int EM_CW = 128;                  // columns of one trailing-update work item
constexpr int EM_RG = EM_WARPS * 16;        // live rows of one trailing-update work item
constexpr size_t EM_SMEM = (size_t)EM_B * EM_WP * 4 + (size_t)EM_B * EM_B * 4;

constexpr uint32_t EM_PIVOT = 0xfffffffeu;   // lead-table mark of a pivot row (ECH_INF = row is zero)

struct EmArgs {
    Fp f;
    uint32_t n_rows, n_cols, ld;    // D' is n_rows x n_cols, row pitch ld (multiple of 4), standard residues
    uint32_t ncp;                   // n_cols rounded up to a multiple of EM_CW
    uint32_t* d;
    const uint32_t* row_nz;         // [n_rows]
    uint32_t* lead;                 // [2][n_rows] double-buffered: lead column | EM_PIVOT | ECH_INF
    uint32_t* live;                 // [n_rows] row | active << 30 | fresh pivot << 31
    uint8_t* xa;                    // [n_rows][NL][32] limbs of X'
    uint8_t* pbt;                   // [NL][ncp][32] limbs of the panel rows, transposed
    uint32_t* piv_row;              // [2][n_rows]: sorted by pivot column | in order of discovery
    uint32_t* piv_col;              // [2][n_rows]
    uint8_t* is_piv_col;            // [n_cols]
    uint32_t* counts;               // [0] rounds, [1] n_piv, [2..3] live rows of the round (double-buffered)
    uint64_t* out_ptr;              // [n_rows + 1] CSR row pointers of the result
};

__device__ __forceinline__ void em_mma_u8(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// old entry + sum_s acc_s * 2^(8s)  mod p
template <int NL>
__device__ __forceinline__ uint32_t em_combine(const int (&acc)[2 * NL - 1][4], int e, uint32_t cin, const Fp& f) {
    if (NL == 2) {
        const uint64_t s = (uint64_t)(uint32_t)acc[0][e] + ((uint64_t)(uint32_t)acc[1][e] << 8) + ((uint64_t)(uint32_t)acc[2][e] << 16);
        return fp_reduce64(s + cin, f);
    } else {
        const uint64_t lo = (uint64_t)(uint32_t)acc[0][e] + ((uint64_t)(uint32_t)acc[1][e] << 8) +
                            ((uint64_t)(uint32_t)acc[2][e] << 16) + ((uint64_t)(uint32_t)acc[3][e] << 24);
        const uint64_t hi = (uint64_t)(uint32_t)acc[4][e] + ((uint64_t)(uint32_t)acc[5][e] << 8) + ((uint64_t)(uint32_t)acc[6][e] << 16);
        return fp_reduce64(lo + cin + ((uint64_t)fp_reduce64(hi, f) << 32), f);
    }
}

template <int NL>
__global__ void __launch_bounds__(EM_THREADS, 1) echelon_mma_kernel(EmArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char em_smem[];
    uint32_t* Wm = reinterpret_cast<uint32_t*>(em_smem);                 // [EM_B][EM_WP]; reused as the B chunk in phase 3
    uint32_t* Tn = Wm + EM_B * EM_WP;                                    // [EM_B][EM_B] monic transformation, row = pivot ordinal
    __shared__ uint32_t sh32[32];
    __shared__ uint32_t s_wcnt[EM_WARPS];
    __shared__ uint32_t s_rsel[EM_B], s_plead[EM_B], s_used[EM_B], s_order[EM_B], s_pcol[EM_B], s_dinv[EM_B];
    __shared__ uint32_t s_urow[EM_B], s_ucol[EM_B];                      // by pivot ordinal: row of D', window-relative column

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nb = gridDim.x, bid = blockIdx.x;
    const uint32_t nr = a.n_rows, nc = a.n_cols, ld = a.ld, p = a.f.p;
    const uint32_t gwarp = bid * EM_WARPS + warp, ngwarps = nb * EM_WARPS;
    const uint32_t lt = (1u << lane) - 1u;

    // ---- leads of the rows, initial states ----------------------------------------------------
    for (uint32_t r = gwarp; r < nr; r += ngwarps) {
        uint32_t l = ECH_INF;
        if (a.row_nz[r]) {
            const uint32_t* row = a.d + (uint64_t)r * ld;
            for (uint32_t j0 = 0; j0 < nc && l == ECH_INF; j0 += 32) {
                const uint32_t j = j0 + lane;
                const uint32_t bal = __ballot_sync(0xffffffffu, j < nc && row[j] != 0);
                if (bal) l = j0 + __ffs(bal) - 1;
            }
        }
        if (lane == 0) a.lead[r] = l;
    }
    for (uint32_t j = bid * EM_THREADS + tid; j < nc; j += nb * EM_THREADS) a.is_piv_col[j] = 0;
    if (bid == 0 && tid == 0) { a.counts[2] = 0; a.counts[3] = 0; }
    grid.sync();

    uint32_t npiv = 0;
    uint32_t* piv_row_u = a.piv_row + nr;   // in order of discovery
    uint32_t* piv_col_u = a.piv_col + nr;
    for (uint32_t round = 0;; ++round) {
        const uint32_t* lc = a.lead + (size_t)(round & 1) * nr;
        uint32_t* ln = a.lead + (size_t)((round + 1) & 1) * nr;

        // ================= phase 1: panel selection and window factorisation (redundant per CTA) ===
        uint32_t mn = ECH_INF;
        for (uint32_t r = tid; r < nr; r += EM_THREADS) { const uint32_t l = __ldcg(lc + r); if (l < EM_PIVOT) mn = min(mn, l); }
        const uint32_t cmin = block_min_u32(mn, sh32);
        if (cmin == ECH_INF) break;
        uint32_t base = 0;
        for (uint32_t r0 = 0; r0 < nr && base < EM_B; r0 += EM_THREADS) {
            const uint32_t r = r0 + tid;
            const uint32_t l = r < nr ? __ldcg(lc + r) : ECH_INF;
            const bool flag = l < EM_PIVOT && l - cmin < (uint32_t)EM_W;
            const uint32_t bal = __ballot_sync(0xffffffffu, flag);
            if (lane == 0) s_wcnt[warp] = __popc(bal);
            __syncthreads();
            uint32_t pre = 0, tot = 0;
            for (uint32_t w = 0; w < EM_WARPS; ++w) { const uint32_t c = s_wcnt[w]; pre += w < warp ? c : 0u; tot += c; }
            const uint32_t pos = base + pre + __popc(bal & lt);
            if (flag && pos < EM_B) { s_rsel[pos] = r; s_plead[pos] = l - cmin; s_used[pos] = 0; }
            base += tot;
            __syncthreads();
        }
        const uint32_t nsel = min(base, (uint32_t)EM_B);
        // window of the selected rows + identity block
        for (uint32_t i = warp; i < nsel; i += EM_WARPS) {
            const uint32_t* row = a.d + (uint64_t)s_rsel[i] * ld;
            uint32_t* w = Wm + i * EM_WP;
            for (uint32_t j = lane; j < EM_W; j += 32) { const uint32_t col = cmin + j; w[j] = col < nc ? __ldcg(row + col) : 0u; }
            w[EM_W + lane] = lane == i ? 1u : 0u;
        }
        __syncthreads();
        uint32_t npv = 0;
        for (; npv < nsel; ++npv) {
            // every warp derives the same pivot: the unused row with the left-most lead
            uint32_t key = ECH_INF;
            if (lane < nsel && !s_used[lane] && s_plead[lane] != ECH_INF) key = (s_plead[lane] << 5) | lane;
            for (int o = 16; o; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
            if (key == ECH_INF) break;
            __syncthreads();        // every warp has read the lead table before any warp rewrites entries of it
            const uint32_t ip = key & 31u, cs = key >> 5;
            const uint32_t* prow = Wm + ip * EM_WP;
            const uint32_t dv = prow[cs];
            for (uint32_t i = warp; i < nsel; i += EM_WARPS) {
                if (i == ip) continue;
                uint32_t* row = Wm + i * EM_WP;
                const uint32_t m = row[cs];
                if (m == 0) continue;
                const uint32_t negm = p - m;
                const bool used = s_used[i] != 0;
                // a pivot row is rescaled from its own pivot on (the pivot row is zero left of cs)
                const uint32_t start = used ? s_pcol[i] : cs + 1;
                for (uint32_t j = start + lane; j < EM_WP; j += 32)
                    row[j] = fp_reduce64((uint64_t)dv * row[j] + (uint64_t)negm * prow[j], a.f);
                __syncwarp();
                if (!used) {   // its lead was cs: find the new one inside the window
                    if (lane == 0) row[cs] = 0;
                    uint32_t nl = ECH_INF;
                    for (uint32_t j0 = cs + 1; j0 < EM_W && nl == ECH_INF; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        const uint32_t bal = __ballot_sync(0xffffffffu, j < EM_W && row[j] != 0);
                        if (bal) nl = j0 + __ffs(bal) - 1;
                    }
                    if (lane == 0) s_plead[i] = nl;
                }
            }
            if (tid == 0) { s_used[ip] = 1; s_order[npv] = ip; s_pcol[ip] = cs; }
            __syncthreads();
        }
        // monic transformation rows, by pivot ordinal
        if (tid < npv) {
            const uint32_t i = s_order[tid];
            s_dinv[tid] = fp_inv32(Wm[i * EM_WP + s_pcol[i]], p);
            s_urow[tid] = s_rsel[i]; s_ucol[tid] = s_pcol[i];
        }
        __syncthreads();
        for (uint32_t x = tid; x < npv * EM_B; x += EM_THREADS) {
            const uint32_t j = x / EM_B, k = x % EM_B;
            Tn[x] = fp_mul(s_dinv[j], Wm[s_order[j] * EM_WP + EM_W + k], a.f);
        }
        __syncthreads();

        // ================= phase 2: update operands =================================================
        const uint32_t c0 = cmin & ~(uint32_t)(EM_CW - 1);
        uint32_t* nlive = a.counts + 2 + (round & 1);
        for (uint32_t r = gwarp; r < nr; r += ngwarps) {
            const uint32_t l = __ldcg(lc + r);
            if (l == ECH_INF) { if (lane == 0) ln[r] = ECH_INF; continue; }      // zero row (from the start, or reduced to zero)
            const bool active = l != EM_PIVOT;
            const uint32_t um = __ballot_sync(0xffffffffu, lane < npv && s_urow[lane] == r);
            uint32_t xv;     // X'[r][lane]
            uint32_t tag;
            if (um) {        // a new pivot row: replaced by T[j] * panel
                const uint32_t j = __ffs(um) - 1;
                xv = Tn[j * EM_B + lane];
                tag = 1u << 31;
                if (lane == 0) {
                    ln[r] = EM_PIVOT;
                    piv_row_u[npiv + j] = r; piv_col_u[npiv + j] = cmin + s_ucol[j]; a.is_piv_col[cmin + s_ucol[j]] = 1;
                }
            } else {
                const uint32_t* row = a.d + (uint64_t)r * ld;
                const uint32_t x = lane < npv ? __ldcg(row + cmin + s_ucol[lane]) : 0u;
                const uint32_t nzx = __ballot_sync(0xffffffffu, x != 0);
                if (nzx == 0) { if (lane == 0) ln[r] = l; continue; }   // untouched by this round
                uint64_t acc = 0;
                for (uint32_t m = nzx; m; m &= m - 1) {
                    const int j = __ffs(m) - 1;
                    acc = fp_mac<false>(acc, __shfl_sync(0xffffffffu, x, j), Tn[j * EM_B + lane], a.f.p2);
                }
                xv = fp_reduce64(acc, a.f);
                xv = xv ? p - xv : 0u;
                tag = active ? 1u << 30 : 0u;
                if (lane == 0) ln[r] = active ? ECH_INF : EM_PIVOT;      // an active row gets its new lead in phase 3
            }
            uint32_t slot = 0;
            if (lane == 0) slot = atomicAdd(nlive, 1u);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (lane == 0) a.live[slot] = r | tag;
#pragma unroll
            for (int li = 0; li < NL; ++li) a.xa[((size_t)slot * NL + li) * 32 + lane] = (uint8_t)(xv >> (8 * li));
        }
        // limbs of the panel rows, transposed: pbt[limb][col][panel row]
        for (uint32_t col = c0 + bid * EM_THREADS + tid; col < a.ncp; col += nb * EM_THREADS) {
            uint32_t w[NL][8];
#pragma unroll
            for (int li = 0; li < NL; ++li)
#pragma unroll
                for (int q = 0; q < 8; ++q) w[li][q] = 0;
            if (col < nc) {
#pragma unroll
                for (uint32_t i = 0; i < EM_B; ++i) {
                    const uint32_t v = i < nsel ? __ldcg(a.d + (uint64_t)s_rsel[i] * ld + col) : 0u;
#pragma unroll
                    for (int li = 0; li < NL; ++li) w[li][i >> 2] |= ((v >> (8 * li)) & 0xffu) << (8 * (i & 3));
                }
            }
#pragma unroll
            for (int li = 0; li < NL; ++li) {
                uint4* dst = reinterpret_cast<uint4*>(a.pbt + ((size_t)li * a.ncp + col) * 32);
                dst[0] = make_uint4(w[li][0], w[li][1], w[li][2], w[li][3]);
                dst[1] = make_uint4(w[li][4], w[li][5], w[li][6], w[li][7]);
            }
        }
        if (bid == 0 && tid == 0) a.counts[2 + ((round + 1) & 1)] = 0;
        npiv += npv;
        grid.sync();

        // ================= phase 3: trailing update on the tensor cores =============================
        {
            const uint32_t nl_rows = __ldcg(nlive);
            const uint32_t nrg = (nl_rows + EM_RG - 1) / EM_RG;
            const uint32_t q0 = c0 / EM_CW, nq = a.ncp / EM_CW - q0;
            const uint32_t g = lane >> 2, t = lane & 3;
            unsigned char* Bs = em_smem;        // [NL][EM_CW][32]
            for (uint32_t item = bid; item < nrg * nq; item += nb) {
                const uint32_t rg = item / nq, q = q0 + item % nq;
                __syncthreads();
                for (uint32_t x = tid; x < NL * EM_CW * 2; x += EM_THREADS) {
                    const uint32_t li = x / (EM_CW * 2), o = x % (EM_CW * 2);
                    reinterpret_cast<uint4*>(Bs)[x] = __ldcg(reinterpret_cast<const uint4*>(a.pbt + ((size_t)li * a.ncp + (size_t)q * EM_CW) * 32) + o);
                }
                __syncthreads();
                const uint32_t sbase = rg * EM_RG + warp * 16;
                if (sbase >= nl_rows) continue;
                const uint32_t slot[2] = {sbase + g, sbase + g + 8};
                uint32_t rowi[2], tag[2];
                uint32_t af[NL][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const bool v = slot[h] < nl_rows;
                    const uint32_t e = v ? __ldcg(a.live + slot[h]) : 0u;
                    rowi[h] = v ? (e & 0x3fffffffu) : ECH_INF; tag[h] = e >> 30;
#pragma unroll
                    for (int li = 0; li < NL; ++li) {
                        uint2 w = make_uint2(0u, 0u);
                        if (v) w = __ldcg(reinterpret_cast<const uint2*>(a.xa + ((size_t)slot[h] * NL + li) * 32) + t);
                        af[li][h] = w.x; af[li][2 + h] = w.y;
                    }
                }
                uint32_t cand[2] = {ECH_INF, ECH_INF};
                for (uint32_t blk = 0; blk < EM_CW / 16; ++blk) {
                    __syncwarp();
                    const uint32_t col = q * EM_CW + blk * 16 + 4 * t;      // this thread's 4 columns
                    int acc[2][2 * NL - 1][4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
#pragma unroll
                        for (int s = 0; s < 2 * NL - 1; ++s) acc[u][s][0] = acc[u][s][1] = acc[u][s][2] = acc[u][s][3] = 0;
                        // fragment column n = g of sub-tile u is chunk column 4*(n/2) + 2u + (n&1): a thread then owns 4 adjacent columns
                        const uint32_t pc = blk * 16 + 4 * (g >> 1) + 2 * u + (g & 1);
                        uint32_t bf[NL][2];
#pragma unroll
                        for (int li = 0; li < NL; ++li) {
                            const uint2 w = *reinterpret_cast<const uint2*>(Bs + ((size_t)li * EM_CW + pc) * 32 + 8 * t);
                            bf[li][0] = w.x; bf[li][1] = w.y;
                        }
#pragma unroll
                        for (int la = 0; la < NL; ++la)
#pragma unroll
                            for (int lb = 0; lb < NL; ++lb) em_mma_u8(acc[u][la + lb], af[la], bf[lb]);
                    }
                    if (col >= ld) continue;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (rowi[h] == ECH_INF) continue;
                        uint4* ptr = reinterpret_cast<uint4*>(a.d + (uint64_t)rowi[h] * ld + col);
                        uint4 cin = make_uint4(0u, 0u, 0u, 0u);
                        if (!(tag[h] & 2u)) cin = *ptr;
                        uint4 o;
                        o.x = em_combine<NL>(acc[0], 2 * h + 0, cin.x, a.f);
                        o.y = em_combine<NL>(acc[0], 2 * h + 1, cin.y, a.f);
                        o.z = em_combine<NL>(acc[1], 2 * h + 0, cin.z, a.f);
                        o.w = em_combine<NL>(acc[1], 2 * h + 1, cin.w, a.f);
                        *ptr = o;
                        if ((tag[h] & 1u) && cand[h] == ECH_INF) {
                            uint32_t c = ECH_INF;
                            if (o.w && col + 3 < nc) c = col + 3;
                            if (o.z && col + 2 < nc) c = col + 2;
                            if (o.y && col + 1 < nc) c = col + 1;
                            if (o.x && col < nc) c = col;
                            cand[h] = c;
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t c = cand[h];
                    c = min(c, __shfl_xor_sync(0xffffffffu, c, 1));
                    c = min(c, __shfl_xor_sync(0xffffffffu, c, 2));
                    if (t == 0 && c != ECH_INF) atomicMin(ln + rowi[h], c);
                }
            }
        }
        grid.sync();
    }

    // ---- pivots sorted by column, sizes of the result rows ----------------------------------------
    if (bid == 0 && tid == 0) a.counts[1] = npiv;
    for (uint32_t k = bid * EM_THREADS + tid; k < npiv; k += nb * EM_THREADS) {
        const uint32_t c = __ldcg(piv_col_u + k);
        uint32_t rank = 0;
        for (uint32_t j = 0; j < npiv; ++j) rank += __ldcg(piv_col_u + j) < c ? 1u : 0u;
        a.piv_col[rank] = c; a.piv_row[rank] = __ldcg(piv_row_u + k);
    }
    grid.sync();
    for (uint32_t s = bid; s < npiv; s += nb) {
        const uint32_t* row = a.d + (uint64_t)a.piv_row[s] * ld;
        const uint32_t c = a.piv_col[s];
        uint32_t n = 0;
        for (uint32_t j = c + tid; j < nc; j += EM_THREADS) n += (j == c || (!a.is_piv_col[j] && row[j])) ? 1u : 0u;
        n = block_sum_u32(n, sh32);
        if (tid == 0) a.out_ptr[s + 1] = n;
    }
    grid.sync();
    if (bid == 0 && tid == 0) {
        uint64_t acc = 0; a.out_ptr[0] = 0;
        for (uint32_t s = 0; s < npiv; ++s) { acc += a.out_ptr[s + 1]; a.out_ptr[s + 1] = acc; }
    }
}

}  // namespace gcu
