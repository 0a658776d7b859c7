// Blocked triangular solve for the multipliers M = -C A^-1, everything GEMM-shaped on the tensor cores
// (role of the reference's closed kernel/reduce_u16|u32.cpp + saxpy.cpp, CMakeLists.txt:81-90).
//
// A (n_top x n_top, unit upper triangular, sparse) is cut into pivot tiles of T columns (T = 128 * 2^j). For tile t
//   cd_t = C_t + M[:, < tT] * A[< tT, tile t]         limb_gemm_tc_kernel<EPI_RMW_LIMBS>  (operand At_t, schur_bt_kernel)
//   M[:, tile t] = -cd_t * Inv_t,  Inv_t = A_tt^-1     limb_gemm_tc_kernel<EPI_LIMBS>      (operand Vt)
// so no row is ever walked window by window: the per-row chain of K/32 dependent steps of round 1's elim_kernel
// (47 % of the device time, flat in the number of GPUs) is gone, and the work is proportional to the rows.
//
// Inv_t for all tiles at once, once per step, by recursive doubling:
//   base    tri_base_kernel: the 128 x 128 diagonal blocks are inverted in shared memory (one CTA each: two levels of
//           the doubling below on the CUDA cores, starting from the inverses of the 32 x 32 windows that staging
//           computes anyway), written as u8 limb planes in both orientations
//             Wr[l][g][c']  = Inv[g][tile start + c']   (row-major:  A-operand of a product)
//             Vt[l][g][k']  = Inv[tile start + k'][g]   (transposed: B-operand of a product)
//           g = global pivot index, c' / k' = position inside the tile: [NL][Kt][T] bytes each, Kt = n_tiles * T
//   level   s -> 2s, all K/(2s) pairs as one batched product each:  inv [[P, Q], [0, R]] = [[P', -P' Q R'], [0, R']]
//             G = P' * Q      A-operand Wr, B-operand Adt (A's diagonal tiles transposed), output limbs Gx
//             H = -(G * R')   A-operand Gx, B-operand Vt, output limbs straight into Wr and (transposed) Vt
// Cost: K T^2 / 3 mod-p multiply-adds (cyclic9-15, K = 40 k, T = 2048: 5.6e10 = 0.2 ms of int8 tensor time).
#pragma once
#include "layout.cuh"

namespace gcu {

constexpr int TRI_BASE = 128;              // diagonal blocks inverted on the CUDA cores

struct TriRowsArgs {
    Layout L;
    const uint32_t* seg;
    const uint2* ent;
    uint32_t* dst;                         // dst[li * ld + (column - col_base)]
    uint32_t ld, width;                    // `width` (multiple of 4) words are zeroed per row before the scatter
    uint32_t t_lo, t_hi;                   // column tiles whose entries are scattered
    uint32_t col_base;                     // block-form column of dst column 0
    uint32_t n_local, n_pad;               // rows li < n_local are scattered, rows li < n_pad zeroed
    uint32_t row_begin, row_step;          // row li is row row_begin + li * row_step of C|D
};

// One CTA per row: zero the row of the dense block, then drop the row's own entries of the tiles [t_lo, t_hi) in.
__global__ void __launch_bounds__(128) tri_rows_kernel(TriRowsArgs a) {
    const uint32_t li = blockIdx.x, tid = threadIdx.x;
    uint32_t* out = a.dst + (uint64_t)li * a.ld;
    for (uint32_t i = tid * 4; i < a.width; i += 128 * 4) *reinterpret_cast<uint4*>(out + i) = make_uint4(0, 0, 0, 0);
    if (li >= a.n_local) return;
    __syncthreads();
    const uint32_t grow = a.L.n_top + a.row_begin + li * a.row_step;
    for (uint32_t t = a.t_lo; t < a.t_hi; ++t) {
        const uint32_t s = a.seg[(uint64_t)t * a.L.n_rows + grow], e = a.seg[(uint64_t)t * a.L.n_rows + grow + 1];
        const uint32_t c0 = lay_tile_start(a.L, t) - a.col_base;
        for (uint32_t k = s + tid; k < e; k += 128) {
            const uint2 en = __ldg(a.ent + k);
            if (en.y) out[c0 + (en.x & 0xffffu)] = en.y;           // padding entries carry value 0 (and the sink column)
        }
    }
}

struct TriBaseArgs {
    Layout L;                              // tile_cols = T
    Fp f;
    const uint32_t* seg;
    const uint2* ent;
    const uint32_t* dinv;                  // inverses of the 32 x 32 diagonal blocks of A (strictly upper part), [n_windows][32][32]
    uint8_t* wr;                           // [nl][kt][T]
    uint8_t* vt;                           // [nl][kt][T]
    uint64_t plane;                        // kt * T
    uint32_t kt, nl;
};

constexpr int TRI_BASE_THREADS = 256;
constexpr int TB_W = TRI_BASE * (TRI_BASE + 1);            // W   [128][129]
constexpr int TB_Q1 = 32 * 33, TB_Q3 = 64 * 65;            // Q1, Q2 [32][33]; Q3, G [64][65]
constexpr size_t TRI_BASE_SMEM = (size_t)(TB_W + 2 * TB_Q1 + 2 * TB_Q3) * sizeof(uint32_t);

// c[i][j] (+)= sum_k a[i][k] b[k][j] mod p for an n x n block, k from k_lo(i) to k_hi(j): delayed reduction in 64 bits
template <bool SMALLP>
__device__ __forceinline__ uint32_t tb_dot(const uint32_t* arow, const uint32_t* bcol, uint32_t bpitch, uint32_t k_lo, uint32_t k_hi, const Fp& f) {
    uint64_t acc = 0;
    for (uint32_t k = k_lo; k < k_hi; ++k) acc = fp_mac<SMALLP>(acc, arow[k], bcol[(size_t)k * bpitch], f.p2);
    return fp_reduce64(acc, f);
}

// One CTA per 128 x 128 diagonal block of A: its inverse W from the inverses of the four 32 x 32 windows (prep_diag_inverse_kernel)
// by two levels of the same doubling the tensor cores continue with: inv [[P, Q], [0, R]] = [[P', -P' Q R'], [0, R']].
// Rows beyond n_top are identity rows. (The first version - back substitution, one barrier per row - took 305 us per step.)
template <bool SMALLP>
__global__ void __launch_bounds__(TRI_BASE_THREADS) tri_base_kernel(TriBaseArgs a) {
    extern __shared__ uint32_t tb_smem[];
    uint32_t* W = tb_smem;
    uint32_t* Q1 = W + TB_W;
    uint32_t* Q2 = Q1 + TB_Q1;
    uint32_t* Q3 = Q2 + TB_Q1;
    uint32_t* G = Q3 + TB_Q3;
    constexpr uint32_t WP = TRI_BASE + 1;
    const uint32_t T = a.L.tile_cols, tid = threadIdx.x;
    const uint32_t g0 = blockIdx.x * TRI_BASE, t = g0 / T, o = g0 - t * T;
    const uint32_t nr = a.L.n_rows, p = a.f.p;
    for (uint32_t i = tid; i < (uint32_t)(TB_W + 2 * TB_Q1 + 2 * TB_Q3); i += TRI_BASE_THREADS) tb_smem[i] = 0;
    __syncthreads();
    // diagonal windows: I + strictly upper part of the window's inverse
    for (uint32_t e = tid; e < 4 * 32 * 32; e += TRI_BASE_THREADS) {
        const uint32_t w = e >> 10, i = (e >> 5) & 31, j = e & 31, win = g0 / WINDOW + w;
        uint32_t v = i == j ? 1u : 0u;
        if (j > i && win < a.L.n_windows) v = a.dinv[(uint64_t)win * WINDOW * WINDOW + i * WINDOW + j];
        W[(32 * w + i) * WP + 32 * w + j] = v;
    }
    // off-diagonal blocks of A from the rows' entries right of their window, inside this 128-block
    if (g0 < a.L.n_top) {      // one contiguous range of ent[]; bits 16-22 of an entry: its row inside the block
        const uint32_t s = a.seg[(uint64_t)t * nr + g0], e = a.seg[(uint64_t)t * nr + min(g0 + TRI_BASE, a.L.n_top)];
        for (uint32_t q = s + tid; q < e; q += TRI_BASE_THREADS) {
            const uint2 en = __ldg(a.ent + q);
            const uint32_t cr = en.x & 0xffffu, i = (en.x >> 16) & 127u;
            if (!en.y || cr >= o + TRI_BASE) continue;
            const uint32_t cb = cr - o;
            if (cb >= 64) { if (i < 64) Q3[i * 65 + cb - 64] = en.y; else Q2[(i - 64) * 33 + cb - 96] = en.y; }
            else Q1[i * 33 + cb - 32] = en.y;
        }
    }
    __syncthreads();
    // level 32 -> 64, two pairs: G = P' Q (P' upper triangular: k >= i), H = -G R' (R' upper triangular: j <= c)
    for (uint32_t e = tid; e < 2 * 32 * 32; e += TRI_BASE_THREADS) {
        const uint32_t pr = e >> 10, i = (e >> 5) & 31, j = e & 31, b = pr * 64;
        const uint32_t* Q = pr ? Q2 : Q1;
        G[(pr * 32 + i) * 65 + j] = tb_dot<SMALLP>(W + (b + i) * WP + b, Q + j, 33, i, 32, a.f);
    }
    __syncthreads();
    for (uint32_t e = tid; e < 2 * 32 * 32; e += TRI_BASE_THREADS) {
        const uint32_t pr = e >> 10, i = (e >> 5) & 31, c = e & 31, b = pr * 64;
        const uint32_t h = tb_dot<SMALLP>(G + (pr * 32 + i) * 65, W + (b + 32) * WP + b + 32 + c, WP, 0, c + 1, a.f);
        W[(b + i) * WP + b + 32 + c] = h ? p - h : 0u;
    }
    __syncthreads();
    // level 64 -> 128
    for (uint32_t e = tid; e < 64 * 64; e += TRI_BASE_THREADS) {
        const uint32_t i = e >> 6, j = e & 63;
        G[i * 65 + j] = tb_dot<SMALLP>(W + i * WP, Q3 + j, 65, i, 64, a.f);
    }
    __syncthreads();
    for (uint32_t e = tid; e < 64 * 64; e += TRI_BASE_THREADS) {
        const uint32_t i = e >> 6, c = e & 63;
        const uint32_t h = tb_dot<SMALLP>(G + i * 65, W + 64 * WP + 64 + c, WP, 0, c + 1, a.f);
        W[i * WP + 64 + c] = h ? p - h : 0u;
    }
    __syncthreads();
    // limb planes, both orientations: 128 consecutive bytes per row
    for (uint32_t l = 0; l < a.nl; ++l) {
        uint8_t* wr = a.wr + l * a.plane + (uint64_t)g0 * T + o;
        uint8_t* vt = a.vt + l * a.plane + (uint64_t)g0 * T + o;
        for (uint32_t e = tid; e < (uint32_t)(TRI_BASE * TRI_BASE); e += TRI_BASE_THREADS) {
            const uint32_t i = e >> 7, c = e & 127;
            wr[(uint64_t)i * T + c] = (uint8_t)(W[i * WP + c] >> (8 * l));
            vt[(uint64_t)i * T + c] = (uint8_t)(W[c * WP + i] >> (8 * l));
        }
    }
}

}  // namespace gcu
