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
//   base    tri_base_kernel: the 128 x 128 diagonal blocks are inverted in shared memory (one CTA each, back
//           substitution over the block's sparse rows), written as u8 limb planes in both orientations
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
    const uint32_t* diag;                  // raw 32 x 32 diagonal blocks of A (strictly upper part), [n_windows][32][32]
    uint8_t* wr;                           // [nl][kt][T]
    uint8_t* vt;                           // [nl][kt][T]
    uint64_t plane;                        // kt * T
    uint32_t kt, nl;
};

// One CTA (128 threads) per 128 x 128 diagonal block: W = (I + N)^-1 by back substitution, rows bottom-up:
// W[i] = e_i - sum_{j > i} N[i][j] W[j]; thread c owns column c. Rows beyond n_top are identity rows.
__global__ void __launch_bounds__(TRI_BASE) tri_base_kernel(TriBaseArgs a) {
    extern __shared__ uint32_t tb_smem[];
    uint32_t* W = tb_smem;                                       // [128][129]
    uint32_t* nrow = tb_smem + TRI_BASE * (TRI_BASE + 1);        // [128]
    const uint32_t T = a.L.tile_cols, c = threadIdx.x;
    const uint32_t g0 = blockIdx.x * TRI_BASE, t = g0 / T, o = g0 - t * T;
    const uint32_t nr = a.L.n_rows, p = a.f.p;
    const uint64_t p2 = a.f.p2;
    for (int i = TRI_BASE - 1; i >= 0; --i) {
        const uint32_t k = g0 + i;
        nrow[c] = 0;
        __syncthreads();
        if (k < a.L.n_top) {
            if (c < 32) {                                        // the row's window of the diagonal blocks
                const uint32_t j = c, wi = k % WINDOW, wb = (k / WINDOW) * WINDOW - g0;
                if (j > wi) { const uint32_t v = a.diag[((uint64_t)(k / WINDOW) * WINDOW + wi) * WINDOW + j]; if (v) nrow[wb + j] = v; }
            }
            const uint32_t s = a.seg[(uint64_t)t * nr + k], e = a.seg[(uint64_t)t * nr + k + 1];
            for (uint32_t q = s + c; q < e; q += TRI_BASE) {     // entries right of the window, inside the tile: ascending columns
                const uint2 en = __ldg(a.ent + q);
                const uint32_t cr = en.x & 0xffffu;
                if (en.y && cr < o + TRI_BASE) nrow[cr - o] = en.y;
            }
        }
        __syncthreads();
        uint32_t w = 0;
        if ((int)c == i) w = 1;
        else if ((int)c > i) {
            uint64_t acc = 0;
            for (uint32_t j = (uint32_t)i + 1; j <= c; ++j) {
                const uint32_t nv = nrow[j];
                if (nv) {
                    if (a.f.small) acc += (uint64_t)nv * W[j * (TRI_BASE + 1) + c];
                    else { acc += (uint64_t)nv * W[j * (TRI_BASE + 1) + c]; if (acc >= p2) acc -= p2; }
                }
            }
            const uint32_t r = fp_reduce64(acc, a.f);
            w = r ? p - r : 0u;
        }
        W[i * (TRI_BASE + 1) + c] = w;
        __syncthreads();
    }
    // limb planes, both orientations: 128 consecutive bytes per row
    for (uint32_t l = 0; l < a.nl; ++l) {
        uint8_t* wr = a.wr + l * a.plane + (uint64_t)g0 * T + o;
        uint8_t* vt = a.vt + l * a.plane + (uint64_t)g0 * T + o;
        for (uint32_t i = 0; i < TRI_BASE; ++i) {
            wr[(uint64_t)i * T + c] = (uint8_t)(W[i * (TRI_BASE + 1) + c] >> (8 * l));
            vt[(uint64_t)i * T + c] = (uint8_t)(W[c * (TRI_BASE + 1) + i] >> (8 * l));
        }
    }
}

}  // namespace gcu
