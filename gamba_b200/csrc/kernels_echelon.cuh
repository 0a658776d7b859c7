// Echelon of the dense D' block (role of the reference's closed
// kernel/echelon_u16|u32 translation units, CMakeLists.txt:81-90) and the
// dense -> sparse compaction of the new rows (open reference code:
// source/matrix.hpp:820-880).
//
// Exact Gauss-Jordan as ONE cooperative kernel, one grid barrier per pivot (a variant clearing 8
// pivots per pass against a snapshotted panel was bit-exact but 1.8x slower on cyclic9-15):
// every CTA re-derives the next pivot (left-most lead among the remaining
// rows) from the double-buffered lead table, so selection needs no barrier of
// its own; the pivot column is never written during an update (it is "dead"
// for all rows but its pivot row afterwards), which removes the barrier between
// reading the multipliers and updating the rows. Pivots come out in ascending
// column order and every pivot column is cleared from ALL other rows, so the
// result is the canonical reduced row echelon form that the host driver (and
// basis.hpp:725-727 in the reference) expects, rows sorted by ascending pivot.
#pragma once
#include <cooperative_groups.h>

#include "layout.cuh"

namespace gcu {
namespace cg = cooperative_groups;

constexpr int ECH_THREADS = 1024;
constexpr uint32_t ECH_INF = 0xffffffffu;

struct EchArgs {
    Fp f;
    uint32_t n_rows, n_cols;        // D' is n_rows x n_cols (non-pivot columns), row-major u32
    uint32_t* d;
    const uint32_t* row_nz;         // [n_rows]
    uint32_t* lead;                 // [n_rows] lead of list row i (owned by one group)
    unsigned long long* best;       // [3] rotating pivot candidates, preset to ~0
    uint32_t* inv_lead;             // [n_rows] inverse of the lead value of list row i
    uint32_t* nz_list;              // [n_rows] rows that are not identically zero
    uint32_t* piv_row;              // [n_rows]
    uint32_t* piv_col;              // [n_rows]
    uint8_t* is_piv_col;            // [n_cols]
    uint32_t* counts;               // [0] n_nz, [1] n_piv
    uint64_t* out_ptr;              // [n_rows + 1] CSR row pointers of the result
};

__device__ __forceinline__ uint32_t block_min_u32(uint32_t v, uint32_t* sh) {
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : ECH_INF;
        for (int o = 16; o; o >>= 1) w = min(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (threadIdx.x == 0) sh[0] = w;
    }
    __syncthreads();
    v = sh[0];
    return v;
}

__device__ __forceinline__ uint64_t block_min_u64(uint64_t v, uint64_t* sh) {
    for (int o = 16; o; o >>= 1) { uint64_t w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint64_t w = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : ~0ull;
        for (int o = 16; o; o >>= 1) { uint64_t z = __shfl_xor_sync(0xffffffffu, w, o); w = z < w ? z : w; }
        if (threadIdx.x == 0) sh[0] = w;
    }
    __syncthreads();
    v = sh[0];
    return v;
}

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t* sh) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0u;
        for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) sh[0] = w;
    }
    __syncthreads();
    v = sh[0];
    return v;
}

constexpr int ECH_Q = 8;                             // columns per thread and super-block (8 x 1024 columns)
constexpr int ECH_OWN = 32;                          // owned rows handled per chunk

// Exact Gauss-Jordan as ONE cooperative kernel, one grid barrier per pivot.
//  * Rows are OWNED by CTAs (row i of the non-zero list belongs to CTA i mod #CTAs); a row's lead
//    lives with its owner, owners publish their smallest remaining lead with one atomicMin into a
//    rotating global slot during the update, and after the grid barrier every CTA reads the next
//    pivot with a single load (no scan, no second barrier).
//  * Inside a CTA the THREADS OWN COLUMNS: thread t holds columns c+1+t+1024q of the pivot row in
//    registers and sweeps them through all owned rows, so every load of a step is independent and
//    in flight together (the row-per-128-threads form paid one ~0.8 us DRAM round trip per 512
//    columns: 47 us per pivot on cyclic9-15, measured with clock64).
//  * The pivot column is never written during an update (it is dead for all rows but its pivot
//    row), which removes the barrier between reading the multipliers and updating the rows.
__global__ void __launch_bounds__(ECH_THREADS, 1) echelon_kernel(EchArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t sh32[32];
    __shared__ uint32_t sh_scan[ECH_THREADS / 32];
    __shared__ uint32_t s_f[ECH_OWN], s_lead[ECH_OWN];
    const uint32_t tid = threadIdx.x, nb = gridDim.x, bid = blockIdx.x;
    const uint32_t nr = a.n_rows, nc = a.n_cols, p = a.f.p;
    unsigned long long* best = a.best;               // [3] rotating (lead << 32 | list index), ~0 = none

    // list of the non-zero rows (block 0, ballot compaction)
    if (bid == 0) {
        uint32_t base = 0;
        for (uint32_t r0 = 0; r0 < nr; r0 += ECH_THREADS) {
            const uint32_t r = r0 + tid;
            const bool nzr = r < nr && a.row_nz[r] != 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, nzr);
            if ((tid & 31) == 0) sh_scan[tid >> 5] = __popc(bal);
            __syncthreads();
            uint32_t pre = base;
            for (uint32_t w = 0; w < (tid >> 5); ++w) pre += sh_scan[w];
            if (nzr) a.nz_list[pre + __popc(bal & ((1u << (tid & 31)) - 1u))] = r;
            uint32_t tot = 0;
            for (uint32_t w = 0; w < ECH_THREADS / 32; ++w) tot += sh_scan[w];
            base += tot;
            __syncthreads();
        }
        if (tid == 0) a.counts[0] = base;
    }
    for (uint32_t j = bid * ECH_THREADS + tid; j < nc; j += nb * ECH_THREADS) a.is_piv_col[j] = 0;
    grid.sync();
    const uint32_t nnz_rows = a.counts[0];

    // leads of the owned rows, their inverses, and the first pivot candidate
    {
        unsigned long long lb = ~0ull;
        for (uint32_t i = bid; i < nnz_rows; i += nb) {
            const uint32_t r = a.nz_list[i];
            const uint32_t* row = a.d + (uint64_t)r * nc;
            uint32_t l = ECH_INF;
            for (uint32_t j0 = 0; j0 < nc && l == ECH_INF; j0 += ECH_THREADS) {
                const uint32_t j = j0 + tid;
                l = block_min_u32(j < nc && row[j] ? j : ECH_INF, sh32);
            }
            if (tid == 0) {
                a.lead[i] = l;
                if (l != ECH_INF) {
                    a.inv_lead[i] = fp_inv_fermat(row[l], a.f);
                    const unsigned long long k = ((unsigned long long)l << 32) | i;
                    lb = k < lb ? k : lb;
                }
            }
        }
        if (tid == 0 && lb != ~0ull) atomicMin(best + 0, lb);
    }
    grid.sync();

    uint32_t step = 0;
    for (;; ++step) {
        const unsigned long long key = *(volatile unsigned long long*)(best + step % 3);
        if (key == ~0ull) break;
        const uint32_t c = (uint32_t)(key >> 32), ip = (uint32_t)key;
        const uint32_t pr = a.nz_list[ip];
        const uint32_t* __restrict__ prow = a.d + (uint64_t)pr * nc;
        const uint32_t inv = a.inv_lead[ip];
        if (bid == 0 && tid == 0) { a.piv_row[step] = pr; a.piv_col[step] = c; a.is_piv_col[c] = 1; best[(step + 2) % 3] = ~0ull; }
        unsigned long long lb = ~0ull;               // thread 0 only
        for (uint32_t i0 = bid; i0 < nnz_rows; i0 += nb * ECH_OWN) {
            // multipliers and leads of (a chunk of) the owned rows
            __syncthreads();
            if (tid < ECH_OWN) {
                const uint32_t i = i0 + tid * nb;
                uint32_t fv = 0, lr = ECH_INF;
                if (i < nnz_rows) {
                    lr = a.lead[i];
                    if (i == ip) { a.lead[i] = ECH_INF; lr = ECH_INF; }
                    else fv = __ldcg(a.d + (uint64_t)a.nz_list[i] * nc + c);
                }
                s_f[tid] = fv; s_lead[tid] = lr;
            }
            __syncthreads();
            for (uint32_t jb = c + 1; jb < nc; jb += ECH_Q * ECH_THREADS) {
                uint32_t pv[ECH_Q];
#pragma unroll
                for (int q = 0; q < ECH_Q; ++q) {
                    const uint32_t j = jb + q * ECH_THREADS + tid;
                    pv[q] = j < nc ? __ldcg(prow + j) : 0u;
                }
                for (uint32_t k = 0; k < ECH_OWN; ++k) {
                    const uint32_t fv = s_f[k];
                    if (fv == 0) continue;              // CTA-uniform
                    uint32_t* __restrict__ row = a.d + (uint64_t)a.nz_list[i0 + k * nb] * nc;
                    const uint32_t mult = p - fp_mul(fv, inv, a.f);
                    uint32_t v[ECH_Q];
#pragma unroll
                    for (int q = 0; q < ECH_Q; ++q) {
                        const uint32_t j = jb + q * ECH_THREADS + tid;
                        v[q] = (j < nc && pv[q]) ? __ldcg(row + j) : 0u;
                    }
#pragma unroll
                    for (int q = 0; q < ECH_Q; ++q) {
                        const uint32_t j = jb + q * ECH_THREADS + tid;
                        if (pv[q]) row[j] = fp_reduce64((uint64_t)v[q] + (uint64_t)mult * pv[q], a.f);
                    }
                }
            }
            // remaining rows whose lead was c: find the new lead (rare: usually one or two rows per pivot)
            for (uint32_t k = 0; k < ECH_OWN; ++k) {
                const uint32_t lr = s_lead[k];
                if (lr == ECH_INF) continue;            // CTA-uniform
                const uint32_t i = i0 + k * nb;
                uint32_t nl = lr;
                if (s_f[k] != 0) {
                    __syncthreads();                    // the row[] stores above
                    const uint32_t* row = a.d + (uint64_t)a.nz_list[i] * nc;
                    nl = ECH_INF;
                    for (uint32_t j0 = c + 1; j0 < nc && nl == ECH_INF; j0 += ECH_THREADS) {
                        const uint32_t j = j0 + tid;
                        nl = block_min_u32(j < nc && __ldcg(row + j) ? j : ECH_INF, sh32);
                    }
                    if (tid == 0) {
                        a.lead[i] = nl;
                        if (nl != ECH_INF) a.inv_lead[i] = fp_inv_fermat(__ldcg(row + nl), a.f);
                    }
                }
                if (tid == 0 && nl != ECH_INF) { const unsigned long long kk = ((unsigned long long)nl << 32) | i; lb = kk < lb ? kk : lb; }
            }
        }
        if (tid == 0 && lb != ~0ull) atomicMin(best + (step + 1) % 3, lb);
        grid.sync();
    }
    const uint32_t npiv = step;
    if (bid == 0 && tid == 0) a.counts[1] = npiv;

    // non-zeros per pivot row over the live columns (its own pivot + non-pivot columns)
    for (uint32_t s = bid; s < npiv; s += nb) {
        const uint32_t* row = a.d + (uint64_t)a.piv_row[s] * nc;
        const uint32_t c = a.piv_col[s];
        uint32_t n = 0;
        for (uint32_t j = c + tid; j < nc; j += ECH_THREADS) n += (j == c || (!a.is_piv_col[j] && row[j])) ? 1u : 0u;
        n = block_sum_u32(n, sh32);
        if (tid == 0) a.out_ptr[s + 1] = n;
    }
    grid.sync();
    if (bid == 0 && tid == 0) {
        uint64_t acc = 0; a.out_ptr[0] = 0;
        for (uint32_t s = 0; s < npiv; ++s) { acc += a.out_ptr[s + 1]; a.out_ptr[s + 1] = acc; }
    }
}

struct EmitArgs {
    Fp f;
    uint32_t n_cols, n_piv;
    const uint32_t* d;
    const uint32_t* piv_row;
    const uint32_t* piv_col;
    const uint8_t* is_piv_col;
    const uint64_t* out_ptr;
    uint32_t* out_col;
    uint32_t* out_val;              // standard or Montgomery residues per f.mont_bits
};

// One CTA per new row: normalise (make monic), drop dead columns, compact in order.
__global__ void __launch_bounds__(256) emit_kernel(EmitArgs a) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    const uint32_t s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* row = a.d + (uint64_t)a.piv_row[s] * a.n_cols;
    const uint32_t c = a.piv_col[s];
    __shared__ uint32_t s_scale;
    if (tid == 0) {
        uint32_t inv = fp_inv32(row[c], a.f.p);
        if (a.f.mont_bits) inv = fp_mul(inv, a.f.r_mod_p, a.f);
        s_scale = inv; s_base = 0;
    }
    __syncthreads();
    const uint32_t scale = s_scale;
    const uint64_t o = a.out_ptr[s];
    for (uint32_t j0 = c; j0 < a.n_cols; j0 += 256) {
        const uint32_t j = j0 + tid;
        uint32_t v = 0;
        if (j < a.n_cols && (j == c || !a.is_piv_col[j])) v = row[j];
        const uint32_t bal = __ballot_sync(0xffffffffu, v != 0);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint32_t pre = s_base;
        for (uint32_t w = 0; w < warp; ++w) pre += s_warp[w];
        if (v) {
            const uint64_t pos = o + pre + __popc(bal & ((1u << lane) - 1u));
            a.out_col[pos] = j;
            a.out_val[pos] = fp_mul(v, scale, a.f);
        }
        __syncthreads();
        if (tid == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
}

}  // namespace gcu
