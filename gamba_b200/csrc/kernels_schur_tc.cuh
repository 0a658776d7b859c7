// Limb-split mod-p products on the 5th-generation tensor cores: tcgen05.mma kind::i8 with the accumulators in
// tensor memory, operands brought in by TMA. Every dense contraction of the engine runs through this kernel:
//   D' = D + M B                                   (the B half of the elimination, "Schur complement")
//   cd_t = C_t + M[:, < tT] A[< tT, tile t]        (blocked triangular solve, kernels_tri.cuh: earlier pivots)
//   M[:, tile t] = -cd_t A_tt^-1                   (blocked triangular solve: the tile's own pivots)
//   G = P^-1 Q, H = -G R^-1                        (recursive doubling of the inverse diagonal tiles, batched)
//   Dc = S D'                                      (Monte-Carlo combinations in front of the echelon)
// Operands are u8 limb planes (2 limbs for p < 2^16, 4 for p < 2^31) with the contraction index contiguous:
// A-operand [NL][rows][k], B-operand [NL][cols][k], both described as 2-D byte tensors.
//
// One CTA (192 threads) per 128 x BN tile of the output (BN = 128 for 2 limbs, 64 for 4); blockIdx.y = batch:
//   warp 0      TMA producer: per k-step of 128 bytes, NL boxes of the A-operand (128 rows x 128 B) and NL boxes
//               of the B-operand (BN rows x 128 B) into a ring of shared-memory stages (SWIZZLE_128B), completion
//               on full[stage]
//   warp 1      allocates the 512 columns of tensor memory; one lane issues, per k-step, 4 (K = 32 each) x NL^2
//               tcgen05.mma  acc[la+lb] += A_la * B_lb^T  (2*NL-1 accumulators of 128 lanes x BN s32 columns);
//               tcgen05.commit releases the stage (empty[stage]) and, at the end of a k-chunk, signals acc_full
//   warps 2-5   epilogue: tcgen05.ld the accumulators (warp w owns lanes 32*(w%4)..), recombine the limb weights,
//               Barrett, then per EPI: read-modify-write the CTA's tile of the u32 output (EPI_RMW), the same plus
//               u8 limb planes of the final value (EPI_RMW_LIMBS), or limb planes only - optionally negated,
//               optionally also transposed, optionally counting the non-zeros (EPI_LIMBS); acc_empty lets the
//               next k-chunk start.
// s32 accumulators are exact for 2^31 / (NL * 255^2) k's: a k-chunk is 128 (NL=2) / 64 (NL=4) k-steps.
// The operands of a 128 x 128 tile are NL * 32 KB per k-step for NL^2 * 2^21 multiply-adds: 64 B/clk/SM at the
// full tensor rate against ~42 B/clk/SM of L2 bandwidth (B300_MICROARCH.md: LTS cap ~6300 B/clk) - the kernel
// is bound by operand delivery from L2, which is what the measured 53-65 % tensor-pipe activity says.
#pragma once
#include <cuda.h>

#include "kernels_schur.cuh"

namespace gcu {

constexpr int TC_BM = 128;
constexpr int TC_BK = 128;                       // bytes of k per stage and row: one 128-byte swizzle atom
constexpr int TC_THREADS = 192;
template <int NL> struct TcCfg;
template <> struct TcCfg<2> { static constexpr int BN = 128, STAGES = 3, KCHUNK = 128; };
template <> struct TcCfg<4> { static constexpr int BN = 64, STAGES = 2, KCHUNK = 64; };
template <int NL> __host__ __device__ constexpr uint32_t tc_stage_bytes() { return (uint32_t)NL * (TC_BM + TcCfg<NL>::BN) * TC_BK; }
template <int NL> __host__ __device__ constexpr size_t tc_smem_bytes() { return (size_t)TcCfg<NL>::STAGES * tc_stage_bytes<NL>() + 1024 /*alignment*/ + 256 /*barriers*/; }

constexpr int EPI_RMW = 0;         // d[row][col] = (d + acc) mod p
constexpr int EPI_RMW_LIMBS = 1;   // the same, and the final value also as u8 limb planes o1 (all rows / columns of the tile: zeros outside)
constexpr int EPI_LIMBS = 2;       // limb planes o1 (and transposed o2) of acc mod p (negated if neg); a single k-chunk

struct TcArgs {
    Fp f;
    uint32_t kp, mp, np;        // rows per limb plane of the A- / B-operand tensor maps (kp: unused by the kernel, kept for the host)
    uint32_t n_rows, n_cols, ld;
    uint32_t* d;
    uint32_t* row_nz;
    uint32_t m_tiles, n_tiles;
    const uint32_t* first_piv;
    uint32_t row_begin, row_step, k_first;
    uint32_t kchunk;            // k-steps between two folds (<= TcCfg<NL>::KCHUNK)
    uint32_t k_end;             // contraction depth in bytes (multiple of 16)
    // operand coordinates: batch b = blockIdx.y; row = l * plane_rows + row0 + b * row_b + (tile row), k = kbase(b) + k-step,
    // kbase(b) = (k0 + b * k_b) mod k_mod (k_mod = 0: no wrap)
    uint32_t a_row0, a_row_b, a_k0, a_k_b;
    uint32_t b_row0, b_row_b, b_k0, b_k_b;
    uint32_t k_mod;
    uint32_t kmin_sub;          // the contraction index starts at this pivot column: first_piv is taken relative to it
    uint32_t skip_hi;           // != 0: a tile whose rows all start at or right of this pivot column has nothing to do (EPI_LIMBS)
    // limb outputs: o1[l * o1_plane + (o1_row0 + b * o1_row_b + row) * o1_ld + ((o1_col0 + b * o1_col_b) mod k_mod) + col]
    //               o2[l * o2_plane + (o2_row0 + b * o2_row_b + col) * o2_ld + ((o2_col0 + b * o2_col_b) mod k_mod) + row]
    uint8_t* o1; uint64_t o1_plane; uint32_t o1_ld, o1_row0, o1_row_b, o1_col0, o1_col_b;
    uint8_t* o2; uint64_t o2_plane; uint32_t o2_ld, o2_row0, o2_row_b, o2_col0, o2_col_b;
    uint32_t tri;               // 1: the A-operand block is upper triangular (rows <= k): k-steps left of the tile's rows are skipped;
                                // 2: the B-operand block is upper triangular as [k][col] (k <= col): k-steps right of the tile's columns are skipped
    // block-sparse B-operand (kernels_schur.cuh, BsArgs): the k-steps of output tile column nt are the occupied blocks
    // klist[kl_base[kl_cb0 + nt] .. kl_base[kl_cb0 + nt + 1]); the operand of block id sits at row id * BN of its limb plane
    const uint32_t* klist;      // NULL: dense operand, k-steps ks_begin .. ks_end
    const uint32_t* kl_base;
    uint32_t kl_cb0;
    uint32_t kl_id0;            // block ids in the operand are relative to this one
    uint32_t neg;               // EPI_LIMBS: store (p - x) mod p
    // EPI_LIMBS: count the non-zero outputs (row < n_rows, col < n_cols): stats[0] += 1, stats[2] += len[col0 + col]
    const uint64_t* cnt_row_ptr; unsigned long long* stats;
};

// ---- PTX wrappers (sm_100a) -------------------------------------------------------------------------------
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "TC_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra TC_DONE_%=;\n\t"
        "bra TC_WAIT_%=;\n\t"
        "TC_DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {      // arrives on `bar` when all tcgen05.mma issued so far have completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {     // 32 lanes x 16 columns of 32 bits
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory operand descriptor: K-major, SWIZZLE_128B, rows of 128 bytes, 8-row atoms 1024 bytes apart
// (cute/arch/mma_sm100_desc.hpp: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout 2 << 61)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t tc_kbase(uint32_t k0, uint32_t k_b, uint32_t b, uint32_t k_mod) {
    const uint32_t k = k0 + b * k_b;
    return k_mod ? k % k_mod : k;
}

// 16 bytes: limb `l` of the 16 values v[0..15]
__device__ __forceinline__ uint4 tc_pack_limb(const uint32_t (&v)[16], int l) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        w[q] = ((v[4 * q] >> (8 * l)) & 0xffu) | (((v[4 * q + 1] >> (8 * l)) & 0xffu) << 8) |
               (((v[4 * q + 2] >> (8 * l)) & 0xffu) << 16) | (((v[4 * q + 3] >> (8 * l)) & 0xffu) << 24);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int NL, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) limb_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                     const __grid_constant__ CUtensorMap map_b, TcArgs a) {
    using C = TcCfg<NL>;
    constexpr int BN = C::BN, STAGES = C::STAGES, NA = 2 * NL - 1;
    constexpr uint32_t STAGE_BYTES = tc_stage_bytes<NL>();
    constexpr uint32_t A_LIMB = TC_BM * TC_BK, B_LIMB = BN * TC_BK;
    // instruction descriptor (kind::i8): D = s32, A = B = u8, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t IDESC = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ uint32_t s_tmem, s_kmin, s_list[2];
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                    // SWIZZLE_128B atoms are 1024-byte aligned
    const uint32_t bars = base + STAGES * STAGE_BYTES;               // full[STAGES], empty[STAGES], acc_full, acc_empty
    auto full_bar = [&](uint32_t s) { return bars + 8 * s; };
    auto empty_bar = [&](uint32_t s) { return bars + 8 * (STAGES + s); };
    const uint32_t acc_full = bars + 8 * (2 * STAGES), acc_empty = bars + 8 * (2 * STAGES + 1);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Heaviest tiles first: the rows to reduce come sorted by descending first pivot, so the last row blocks contract over the most
    // k-steps. With more tiles than SMs the light ones then fill in behind the heavy ones (before: every wave waited for its
    // heaviest tile; ncu: SMs busy 55 % of such a launch).
    const uint32_t mt = a.m_tiles - 1u - blockIdx.x / a.n_tiles, nt = blockIdx.x % a.n_tiles, bb = blockIdx.y;
    const uint32_t row0 = mt * TC_BM, col0 = nt * BN;

    if (tid == 0) {
        for (uint32_t s = 0; s < (uint32_t)STAGES; ++s) { tc_mbar_init(full_bar(s), 1); tc_mbar_init(empty_bar(s), 1); }
        tc_mbar_init(acc_full, 1); tc_mbar_init(acc_empty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_kmin = 0xffffffffu;
    }
    if (warp == 1) {                                                 // 512 columns of tensor memory for this CTA
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // first pivot column any row of the tile has: everything left of it is zero in the A-operand
    if (tid < (uint32_t)TC_BM) {
        const uint32_t li = row0 + tid;
        if (li < a.n_rows) atomicMin(&s_kmin, a.first_piv ? a.first_piv[a.row_begin + li * a.row_step] : a.k_first);
    }
    __syncthreads();
    const bool skip_all = EPI == EPI_LIMBS && a.skip_hi && s_kmin >= a.skip_hi;   // the output stays what the host zeroed
    uint32_t ks_end = (a.k_end + TC_BK - 1) / TC_BK;
    if (a.tri == 2) ks_end = min(ks_end, (col0 + (uint32_t)BN + TC_BK - 1) / TC_BK);
    uint32_t ks_begin = (s_kmin > a.kmin_sub ? s_kmin - a.kmin_sub : 0u) / TC_BK;
    if (a.tri == 1) ks_begin = max(ks_begin, row0 / TC_BK);
    ks_begin = min(ks_begin, ks_end);
    if (a.klist) {            // block-sparse: the list of this column block, from the first k-block any row of the tile reaches
        if (tid == 0) {
            uint32_t lo = a.kl_base[a.kl_cb0 + nt];
            const uint32_t hi = a.kl_base[a.kl_cb0 + nt + 1];
            uint32_t l = lo, h = hi;
            while (l < h) { const uint32_t m = (l + h) >> 1; if (a.klist[m] < ks_begin) l = m + 1; else h = m; }
            lo = l; l = lo; h = hi;
            while (l < h) { const uint32_t m = (l + h) >> 1; if (a.klist[m] < ks_end) l = m + 1; else h = m; }
            s_list[0] = lo; s_list[1] = l;
        }
        __syncthreads();
    }
    const uint32_t l_lo = a.klist ? s_list[0] : 0u;
    const uint32_t n_ks = skip_all ? 0u : a.klist ? s_list[1] - l_lo : ks_end - ks_begin;
    const uint32_t n_chunks = n_ks ? (n_ks + a.kchunk - 1) / a.kchunk : 0;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const int32_t ka = (int32_t)tc_kbase(a.a_k0, a.a_k_b, bb, a.k_mod), kb = (int32_t)tc_kbase(a.b_k0, a.b_k_b, bb, a.k_mod);
            const int32_t ra = (int32_t)(a.a_row0 + bb * a.a_row_b + row0), rb = (int32_t)(a.b_row0 + bb * a.b_row_b + col0);
            for (uint32_t i = 0; i < n_ks; ++i) {
                const uint32_t s = i % STAGES, ph = (i / STAGES) & 1;
                tc_mbar_wait(empty_bar(s), ph ^ 1);
                tc_mbar_expect_tx(full_bar(s), STAGE_BYTES);
                const uint32_t dst = base + s * STAGE_BYTES;
                const int32_t k0 = (int32_t)((a.klist ? a.klist[l_lo + i] : ks_begin + i) * TC_BK);
#pragma unroll
                for (int l = 0; l < NL; ++l) tc_tma_load_2d(dst + l * A_LIMB, &map_a, full_bar(s), ka + k0, (int32_t)(l * a.mp) + ra);
                if (a.klist) {
#pragma unroll
                    for (int l = 0; l < NL; ++l) tc_tma_load_2d(dst + NL * A_LIMB + l * B_LIMB, &map_b, full_bar(s), 0, (int32_t)(l * a.np + (l_lo + i - a.kl_id0) * BN));
                } else {
#pragma unroll
                    for (int l = 0; l < NL; ++l) tc_tma_load_2d(dst + NL * A_LIMB + l * B_LIMB, &map_b, full_bar(s), kb + k0, (int32_t)(l * a.np) + rb);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (uint32_t i = 0; i < n_ks; ++i) {
                const uint32_t s = i % STAGES, ph = (i / STAGES) & 1;
                const uint32_t in_chunk = i % a.kchunk, chunk = i / a.kchunk;
                if (in_chunk == 0 && chunk > 0) { tc_mbar_wait(acc_empty, (chunk - 1) & 1); tc_fence_after(); }   // epilogue drained the accumulators
                tc_mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = base + s * STAGE_BYTES, sb = sa + NL * A_LIMB;
#pragma unroll
                for (int kk = 0; kk < TC_BK / 32; ++kk)
#pragma unroll
                    for (int la = 0; la < NL; ++la)
#pragma unroll
                        for (int lb = 0; lb < NL; ++lb) {
                            // the first product into an accumulator of a k-chunk overwrites it
                            const uint32_t first = (in_chunk == 0 && kk == 0 && (la == 0 || lb == NL - 1)) ? 1u : 0u;
                            tc_mma_i8(tmem + (uint32_t)(la + lb) * BN, tc_smem_desc(sa + la * A_LIMB + kk * 32),
                                      tc_smem_desc(sb + lb * B_LIMB + kk * 32), IDESC, first ^ 1u);
                        }
                tc_commit(empty_bar(s));                             // the stage is free once these products have read it
                if (in_chunk + 1 == a.kchunk || i + 1 == n_ks) tc_commit(acc_full);
            }
        }
    } else {
        // ===== epilogue: warps 2..5, warp w owns the tensor-memory lanes 32 * (w % 4) .. + 31 =====
        const uint32_t quarter = warp & 3u;
        const uint32_t trow = quarter * 32 + lane;                   // row inside the tile
        const uint32_t row = row0 + trow;
        const bool row_ok = row < a.n_rows;
        uint32_t* drow = EPI == EPI_LIMBS ? nullptr : a.d + (uint64_t)(row_ok ? row : 0) * a.ld;
        uint8_t* o1p = nullptr; uint8_t* o2p = nullptr;
        if (EPI != EPI_RMW) {
            if (a.o1) o1p = a.o1 + (uint64_t)(a.o1_row0 + bb * a.o1_row_b + row) * a.o1_ld + tc_kbase(a.o1_col0, a.o1_col_b, bb, a.k_mod) + col0;
            if (a.o2) o2p = a.o2 + (uint64_t)(a.o2_row0 + bb * a.o2_row_b + col0) * a.o2_ld + tc_kbase(a.o2_col0, a.o2_col_b, bb, a.k_mod) + row;
        }
        uint32_t any = 0;
        unsigned long long cnt_piv = 0, cnt_fma = 0;
        for (uint32_t chunk = 0; chunk < n_chunks; ++chunk) {
            tc_mbar_wait(acc_full, chunk & 1);
            tc_fence_after();
            const bool last = chunk + 1 == n_chunks;
            for (uint32_t c0 = 0; c0 < (uint32_t)BN; c0 += 16) {
                uint32_t v[NA][16];
#pragma unroll
                for (int s = 0; s < NA; ++s) tc_ld16(tmem + ((quarter * 32u) << 16) + (uint32_t)s * BN + c0, v[s]);
                tc_ld_wait();
                const uint32_t col = col0 + c0;
                uint32_t o[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    uint64_t sum;
                    if (NL == 2) {
                        sum = (uint64_t)v[0][e] + ((uint64_t)v[1][e] << 8) + ((uint64_t)v[2][e] << 16);
                    } else {
                        const uint64_t lo = (uint64_t)v[0][e] + ((uint64_t)v[1][e] << 8) + ((uint64_t)v[2][e] << 16) + ((uint64_t)v[3][e] << 24);
                        const uint64_t hi = (uint64_t)v[4 % NA][e] + ((uint64_t)v[5 % NA][e] << 8) + ((uint64_t)v[6 % NA][e] << 16);
                        sum = lo + ((uint64_t)fp_reduce64(hi, a.f) << 32);
                    }
                    if (EPI == EPI_LIMBS) {
                        const uint32_t x = fp_reduce64(sum, a.f);
                        o[e] = (a.neg && x) ? a.f.p - x : x;
                    } else {
                        o[e] = fp_reduce64(sum, a.f);                                          // the old entry is added below
                    }
                }
                if (EPI != EPI_LIMBS) {
                    const bool in = row_ok && col < a.n_cols;
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        if (in && col + e < a.n_cols) {
                            uint32_t* ptr = drow + col + e;
                            const uint4 cin = *reinterpret_cast<const uint4*>(ptr);       // ld is a multiple of 4: the vector stays inside the row
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t ci = j == 0 ? cin.x : j == 1 ? cin.y : j == 2 ? cin.z : cin.w;
                                uint32_t r = o[e + j] + ci;                                  // both < p < 2^31
                                if (r >= a.f.p) r -= a.f.p;
                                o[e + j] = col + e + j < a.n_cols ? r : ci;                  // padding columns of the pitch keep their content
                            }
                            *reinterpret_cast<uint4*>(ptr) = make_uint4(o[e], o[e + 1], o[e + 2], o[e + 3]);
                            if (last) any |= o[e] | o[e + 1] | o[e + 2] | o[e + 3];
                            if (EPI == EPI_RMW_LIMBS) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) if (col + e + j >= a.n_cols) o[e + j] = 0;
                            }
                        } else if (EPI == EPI_RMW_LIMBS) {
                            o[e] = o[e + 1] = o[e + 2] = o[e + 3] = 0;                       // outside the output: zero limbs
                        }
                    }
                }
                if (EPI != EPI_RMW && (EPI == EPI_LIMBS || last)) {
                    if (o1p) {
#pragma unroll
                        for (int l = 0; l < NL; ++l) *reinterpret_cast<uint4*>(o1p + l * a.o1_plane + c0) = tc_pack_limb(o, l);
                    }
                    if (o2p) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
#pragma unroll
                            for (int l = 0; l < NL; ++l) o2p[l * a.o2_plane + (uint64_t)(c0 + e) * a.o2_ld] = (uint8_t)(o[e] >> (8 * l));
                    }
                    if (EPI == EPI_LIMBS && a.stats) {      // lane e counts the non-zeros of column col + e over the warp's rows
                        uint32_t mine = 0;
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const uint32_t m = __ballot_sync(0xffffffffu, o[e] != 0 && row_ok);
                            if ((int)lane == e) mine = __popc(m);
                        }
                        if (mine && col + lane < a.n_cols) {
                            const uint32_t k = a.o1_col0 + col + lane;
                            cnt_piv += mine; cnt_fma += (unsigned long long)mine * (a.cnt_row_ptr[k + 1] - a.cnt_row_ptr[k]);
                        }
                    }
                }
            }
            tc_fence_before();
            tc_mbar_arrive(acc_empty);
        }
        if (EPI == EPI_RMW && n_chunks == 0 && row_ok && a.row_nz)                 // nothing to add: the flag comes from D alone
            for (uint32_t c = col0; c < min(col0 + (uint32_t)BN, a.n_cols); ++c) any |= drow[c];
        if (EPI == EPI_RMW_LIMBS && n_chunks == 0 && o1p) {                        // nothing to add: the limbs of what is there
            for (uint32_t c0 = 0; c0 < (uint32_t)BN; c0 += 16) {
                uint32_t o[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) o[e] = (row_ok && col0 + c0 + e < a.n_cols) ? drow[col0 + c0 + e] : 0u;
#pragma unroll
                for (int l = 0; l < NL; ++l) *reinterpret_cast<uint4*>(o1p + l * a.o1_plane + c0) = tc_pack_limb(o, l);
            }
        }
        if (row_ok && any && a.row_nz) a.row_nz[row] = 1u;
        if (EPI == EPI_LIMBS && a.stats) {
            for (int o_ = 16; o_; o_ >>= 1) { cnt_piv += __shfl_xor_sync(0xffffffffu, cnt_piv, o_); cnt_fma += __shfl_xor_sync(0xffffffffu, cnt_fma, o_); }
            if (lane == 0 && cnt_piv) { atomicAdd(a.stats + 0, cnt_piv); atomicAdd(a.stats + 2, cnt_fma); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- microbenchmark: the int8 tensor rate one SM sustains with its operands already in shared memory -----------------
// (the denominator of bench.py's tensor roofline: MEASURED_PEAKS.json only has cuBLAS bf16). One CTA per SM, one thread
// issues `iters` x 4 tcgen05.mma kind::i8 of 128 x N x 32 into one accumulator, back to back, and waits for the last commit.
__host__ __device__ constexpr size_t tc_peak_smem(uint32_t n) { return (size_t)(TC_BM + n) * TC_BK + 1024 + 64; }

template <int N>
__global__ void __launch_bounds__(128, 1) tc_peak_kernel(uint32_t iters) {
    constexpr uint32_t IDESC = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    extern __shared__ unsigned char tcp_raw[];
    __shared__ uint32_t s_tmem;
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(tcp_raw);
    const uint32_t base = (raw + 1023u) & ~1023u, bar = base + (TC_BM + N) * TC_BK;
    for (uint32_t i = threadIdx.x; i < (uint32_t)(TC_BM + N) * TC_BK / 4; i += 128) reinterpret_cast<uint32_t*>(tcp_raw + (base - raw))[i] = i * 2654435761u;
    if (threadIdx.x == 0) { tc_mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the generic-proxy fill above is read by the tensor cores
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        for (uint32_t it = 0; it < iters; ++it)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                tc_mma_i8(tmem, tc_smem_desc(base + kk * 32), tc_smem_desc(base + TC_BM * TC_BK + kk * 32), IDESC, (it | kk) ? 1u : 0u);
        tc_commit(bar);
        tc_mbar_wait(bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace gcu
