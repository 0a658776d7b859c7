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
//   Mx[limb][row][k]   u8 limbs of the multipliers, written by elim_kernel as it finds them
//   Bt[limb][col][k]   u8 limbs of B, transposed (k contiguous), scattered once per step by
//                      schur_bt_kernel from the staged entries of the non-pivot tiles
//   D'[row][col]       u32 residues; holds D when the product starts
//
// schur_mma_kernel: one CTA per 128 x 128 tile of D' (NL = 2 limbs, p < 2^16) or 64 x 64 (NL = 4,
// p < 2^31), k-steps of 64 staged by cp.async into XOR-swizzled shared memory (4 / 3 stages),
// ldmatrix fragments, mma.sync.m16n8k32 u8*u8->s32 into 2*NL-1 accumulators per output (limb products
// of equal weight share one). s32 accumulators take 2^31 / (NL * 255^2) k's exactly, so every
// `kchunk` k-steps the accumulators are folded into D' (read-modify-write of the CTA's own tile,
// Barrett reduction) and cleared. k-steps left of the first pivot column of the tile's rows are skipped.
//
// Tile order: blockIdx.x = n_tile * m_tiles + m_tile, so CTAs resident together share their Bt tile
// (read from HBM once, served to the other row blocks by L2).
#pragma once
#include "kernels_echelon_mma.cuh"
#include "kernels_elim.cuh"
#include "layout.cuh"

namespace gcu {

constexpr int SC_BK = 64;                  // bytes of k per stage and row

template <int NL> struct SchurCfg;
template <> struct SchurCfg<2> {           // p < 2^16
    static constexpr int WARPS_M = 2, WARPS_N = 4, WM = 64, WN = 32, STAGES = 4, MIN_CTAS = 1;   // 192 accumulator registers
};
template <> struct SchurCfg<4> {           // p < 2^31
    static constexpr int WARPS_M = 2, WARPS_N = 4, WM = 32, WN = 16, STAGES = 4, MIN_CTAS = 1;   // 112 accumulator registers
};
template <int NL> __host__ __device__ constexpr int schur_bm() { return SchurCfg<NL>::WARPS_M * SchurCfg<NL>::WM; }
template <int NL> __host__ __device__ constexpr int schur_bn() { return SchurCfg<NL>::WARPS_N * SchurCfg<NL>::WN; }
template <int NL> __host__ __device__ constexpr int schur_threads() { return SchurCfg<NL>::WARPS_M * SchurCfg<NL>::WARPS_N * 32; }
template <int NL> __host__ __device__ constexpr size_t schur_smem() { return (size_t)SchurCfg<NL>::STAGES * NL * (schur_bm<NL>() + schur_bn<NL>()) * SC_BK; }
template <int NL> __host__ __device__ constexpr uint32_t schur_kchunk() { return NL == 2 ? 256u : 128u; }   // k-steps: NL * 255^2 * 64 * kchunk < 2^31

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
};

// One warp per pivot row k: its entries in the non-pivot tiles become column k of Bt. Consecutive
// warps write adjacent bytes of the same sectors, which L2 merges before they reach HBM.
__global__ void __launch_bounds__(256) schur_bt_kernel(BtArgs a) {
    const Layout& L = a.L;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t k = blockIdx.x * 8 + warp;
    if (k >= L.n_top) return;
    const uint32_t nr = L.n_rows;
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

struct SchurArgs {
    Fp f;
    const uint8_t* mx;          // [NL][mp][kp]
    const uint8_t* bt;          // [NL][np][kp]
    uint64_t mx_plane, bt_plane;
    uint32_t kp;                // n_top rounded up to SC_BK
    uint32_t n_rows, n_cols, ld;
    uint32_t* d;                // D' block of this rank
    uint32_t* row_nz;           // set to 1 for rows with a non-zero entry
    uint32_t m_tiles, n_tiles;
    const uint32_t* first_piv;  // first pivot column per row to reduce (NULL: k_first for all rows)
    uint32_t row_begin, row_step, k_first;
    uint32_t kchunk;            // k-steps between two folds of the accumulators (<= schur_kchunk<NL>())
};

__device__ __forceinline__ void sc_ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void sc_cp_async16(uint32_t saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}

template <int NL>
__global__ void __launch_bounds__(schur_threads<NL>(), SchurCfg<NL>::MIN_CTAS) schur_mma_kernel(SchurArgs a) {
    using C = SchurCfg<NL>;
    constexpr int BM = schur_bm<NL>(), BN = schur_bn<NL>(), THREADS = schur_threads<NL>();
    constexpr int NA = 2 * NL - 1, MI = C::WM / 16, NI = C::WN / 8, STAGES = C::STAGES;
    constexpr uint32_t STAGE_BYTES = NL * (BM + BN) * SC_BK;
    constexpr int ROWS_PER_PASS = THREADS / 4;                  // smem rows filled by one pass of the CTA
    constexpr int PASSES = NL * (BM + BN) / ROWS_PER_PASS;
    static_assert(BM % ROWS_PER_PASS == 0 || ROWS_PER_PASS % BM == 0, "fill passes must not straddle operands");
    extern __shared__ __align__(128) unsigned char sc_smem[];
    __shared__ uint32_t s_kmin;
    const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sc_smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t wm = warp % C::WARPS_M, wn = warp / C::WARPS_M;
    const uint32_t g = lane >> 2, tig = lane & 3;
    const uint32_t mt = blockIdx.x % a.m_tiles, nt = blockIdx.x / a.m_tiles;
    const uint32_t row0 = mt * BM, col0 = nt * BN;

    // first pivot column any row of the tile has: everything left of it is zero in Mx
    if (tid == 0) s_kmin = 0xffffffffu;
    __syncthreads();
    for (uint32_t r = tid; r < (uint32_t)BM; r += THREADS) {
        const uint32_t li = row0 + r;
        if (li < a.n_rows) atomicMin(&s_kmin, a.first_piv ? a.first_piv[a.row_begin + li * a.row_step] : a.k_first);
    }
    __syncthreads();
    const uint32_t ks_end = a.kp / SC_BK;
    const uint32_t ks_begin = min(s_kmin / SC_BK, ks_end);

    // fill: pass j of the CTA copies smem rows [j * ROWS_PER_PASS, (j+1) * ROWS_PER_PASS), 4 chunks of 16 B each
    const uint32_t fr = tid >> 2, fc = tid & 3;
    const uint8_t* src[PASSES];
#pragma unroll
    for (int j = 0; j < PASSES; ++j) {
        const uint32_t rr = j * ROWS_PER_PASS + fr;             // smem row: [NL][BM] of Mx, then [NL][BN] of Bt
        if (rr < (uint32_t)(NL * BM)) src[j] = a.mx + (uint64_t)(rr / BM) * a.mx_plane + (uint64_t)(row0 + rr % BM) * a.kp + fc * 16;
        else { const uint32_t q = rr - NL * BM; src[j] = a.bt + (uint64_t)(q / BN) * a.bt_plane + (uint64_t)(col0 + q % BN) * a.kp + fc * 16; }
    }
    const uint32_t fill_dst = fr * SC_BK + ((fc ^ ((fr >> 1) & 3u)) << 4);
    auto load_stage = [&](uint32_t stage, uint32_t ks) {
        const uint32_t dst = sm_base + stage * STAGE_BYTES + fill_dst;
#pragma unroll
        for (int j = 0; j < PASSES; ++j) sc_cp_async16(dst + j * ROWS_PER_PASS * SC_BK, src[j] + (uint64_t)ks * SC_BK);
    };

    // ldmatrix sources of this lane: byte offset of its row inside a stage + the swizzled 16-byte chunk for
    // each half (kk) of the k-step. Fragment / limb offsets are multiples of 8 rows, which leave the swizzle alone.
    const uint32_t a_row = wm * C::WM + (lane & 7) + ((lane >> 3) & 1) * 8;                    // + mi * 16 + l * BM
    const uint32_t b_row = NL * BM + wn * C::WN + (lane & 7) + (lane >> 4) * 8;               // + nj * 16 + l * BN
    uint32_t a_off[2], b_off[2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        a_off[kk] = a_row * SC_BK + (((kk * 2 + (lane >> 4)) ^ ((a_row >> 1) & 3u)) << 4);
        b_off[kk] = b_row * SC_BK + (((kk * 2 + ((lane >> 3) & 1)) ^ ((b_row >> 1) & 3u)) << 4);
    }

    int acc[MI][NI][NA][4];
    auto clear = [&]() {
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                for (int s = 0; s < NA; ++s) acc[mi][ni][s][0] = acc[mi][ni][s][1] = acc[mi][ni][s][2] = acc[mi][ni][s][3] = 0;
    };
    // D'[tile] += sum_s acc_s 2^(8s)  (mod p); the CTA owns its tile
    auto fold = [&](bool last) {
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t row = row0 + wm * C::WM + mi * 16 + g + h * 8;
                if (row >= a.n_rows) continue;
                uint32_t any = 0;
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) {
                    const uint32_t col = col0 + wn * C::WN + ni * 8 + 2 * tig;
                    if (col >= a.n_cols) continue;
                    uint32_t* ptr = a.d + (uint64_t)row * a.ld + col;
                    if (col + 1 < a.n_cols) {
                        const uint2 cin = *reinterpret_cast<const uint2*>(ptr);
                        uint2 o;
                        o.x = em_combine<NL>(acc[mi][ni], 2 * h + 0, cin.x, a.f);
                        o.y = em_combine<NL>(acc[mi][ni], 2 * h + 1, cin.y, a.f);
                        *reinterpret_cast<uint2*>(ptr) = o;
                        any |= o.x | o.y;
                    } else {
                        const uint32_t o = em_combine<NL>(acc[mi][ni], 2 * h + 0, *ptr, a.f);
                        *ptr = o;
                        any |= o;
                    }
                }
                if (last && any) a.row_nz[row] = 1u;
            }
    };

    clear();
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (ks_begin + s < ks_end) load_stage(s, ks_begin + s);
        cp_async_commit();
    }
    uint32_t since_fold = 0;
    for (uint32_t ks = ks_begin; ks < ks_end; ++ks) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const uint32_t nxt = ks + STAGES - 1;
        if (nxt < ks_end) load_stage((nxt - ks_begin) % STAGES, nxt);
        cp_async_commit();
        const uint32_t st = sm_base + ((ks - ks_begin) % STAGES) * STAGE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            uint32_t bf[NI][NL][2];
#pragma unroll
            for (int nj = 0; nj < NI / 2; ++nj)
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    uint32_t r[4];
                    sc_ldmatrix_x4(r, st + b_off[kk] + (nj * 16 + l * BN) * SC_BK);
                    bf[2 * nj][l][0] = r[0]; bf[2 * nj][l][1] = r[1]; bf[2 * nj + 1][l][0] = r[2]; bf[2 * nj + 1][l][1] = r[3];
                }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int la = 0; la < NL; ++la) {
                    uint32_t af[4];
                    sc_ldmatrix_x4(af, st + a_off[kk] + (mi * 16 + la * BM) * SC_BK);
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                        for (int lb = 0; lb < NL; ++lb) em_mma_u8(acc[mi][ni][la + lb], af, bf[ni][lb]);
                }
        }
        if (++since_fold == a.kchunk && ks + 1 < ks_end) { fold(false); clear(); since_fold = 0; }
    }
    cp_async_wait<0>();
    fold(true);
}

}  // namespace gcu
