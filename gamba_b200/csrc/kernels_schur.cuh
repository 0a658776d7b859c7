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
//   Mx[limb][row][k]   u8 limbs of the multipliers, written by the solve (kernels_tri.cuh / elim_kernel) as it finds them
//   Bt[limb][col][k]   u8 limbs of B, transposed (k contiguous), built once per step by
//                      operand_build_kernel from the staged entries of the non-pivot tiles
//   D'[row][col]       u32 residues; holds D when the product starts
//
// The product itself is limb_gemm_tc_kernel (kernels_schur_tc.cuh: tcgen05 + TMA); this file holds the kernels that
// build its operands. (Round 1's mma.sync version of the product was removed: one product kernel, no fallback.)
#pragma once
#include "kernels_echelon_mma.cuh"
#include "kernels_elim.cuh"
#include "layout.cuh"

namespace gcu {

constexpr int SC_BK = 64;                  // bytes of k per stage and row


// Operand builder: the entries of the pivot rows, transposed into dense u8 limb planes op[l][column][k] (k contiguous).
// One CTA per (block of 128 pivot rows, chunk of OP_CH<NL> columns of one column tile): the block is assembled in shared
// memory from the rows' entries of that tile (tile-major ent[]: short, contiguous segments) and written out as whole
// 128-byte rows - every byte of the operand is written exactly once, zeros included, so there is no memset and no
// scattered byte store in HBM (round 1's per-entry scatter: 2.0 ms + 0.3 ms of memsets per cyclic9-15 step, now ~0.5 ms).
//   mode OP_B    columns = the non-pivot tiles: Bt[l][col][k], the right operand of D' = D + M B
//   mode OP_AT   blockIdx.z = pivot tile t >= 1, pivot rows of the EARLIER tiles: At_t[l][col in tile t][k < t T]
//   mode OP_ADT  blockIdx.z = pivot tile t, its OWN pivot rows (entries right of the row's window):
//                Adt[l][t T + col][k - t T]  (kernels_tri.cuh: the Q blocks of the recursive doubling)
constexpr int OP_B = 0, OP_AT = 1, OP_ADT = 2;
constexpr int OP_KB = 128;                 // pivot rows per block = bytes of one output row segment
constexpr int OP_PITCH = OP_KB + 16;       // shared-memory row pitch (16-byte aligned, spreads the banks)
template <int NL> __host__ __device__ constexpr int op_ch() { return 512 / NL; }          // columns per chunk
template <int NL> __host__ __device__ constexpr size_t op_smem() { return (size_t)NL * op_ch<NL>() * OP_PITCH; }

struct OpArgs {
    Layout L;
    const uint32_t* seg;
    const uint2* ent;
    uint8_t* out;
    uint64_t plane;             // bytes per limb plane (OP_B, OP_ADT)
    uint32_t pitch;             // bytes per operand row (OP_B: kp, OP_ADT: tile_cols)
    uint32_t rows;              // OP_B: operand rows (non-pivot columns padded to 128)
    uint32_t mode;
    const uint64_t* at_off;     // OP_AT: byte offset of tile t's operand
    // block-sparse operands (OP_AT, OP_B): only the blocks of 128 pivot rows x cbw columns that hold an entry are stored,
    // out[l][block id][column in block][128 k bytes]; bitmap / rank / blk_base: BsArgs below
    const uint32_t* bitmap;     // NULL: dense operand
    const uint32_t* rank;
    const uint32_t* blk_base;
    uint32_t kb_words, cbw, n_cb_piv;
    uint32_t z0;                // added to blockIdx.z (OP_AT for one pivot tile at a time)
    uint32_t id0;               // block-sparse: block ids are stored relative to this one (the operand of ONE pivot tile, or of B alone)
};

// ---- block-sparse operands ------------------------------------------------------------------------------------------
// A and B of a katsura / eco step hold 0.1-0.5 % non-zeros, and at the granularity of the product kernel's operand boxes
// (128 pivot rows x 128 columns; 64 columns for 4 limbs) only 12-30 % of the blocks hold an entry at all
// (profiles/r02_block_occupancy.txt: kat15-15, K = 440-517 k: 16-19 %; cyclic9-15: 88 %). The products M * A and M * B then
// skip the empty blocks: per column block the list of occupied k-blocks, the operand stored block-compressed.
//   bitmap[cb][kb / 32]  bit kb % 32: block (k-block kb, column block cb) holds an entry; pivot column blocks first
//                        (cb = column / cbw over the tile-padded pivot range), then the non-pivot ones
//   rank[cb][w]          occupied blocks of column block cb below bitmap word w
//   blk_base[cb]         first block id of column block cb (exclusive scan of the counts; blk_base[n_cb] = blocks in total)
//   klist[id]            k-block of block id: ascending inside a column block
struct BsArgs {
    Layout L;
    const uint32_t* seg;
    const uint2* ent;
    uint32_t cbw, kb_words, n_cb_piv, n_cb;
    uint32_t* bitmap;
    uint32_t* rank;
    uint32_t* blk_base;         // [n_cb + 1]: counts before the scan, bases after
    uint32_t* klist;
    uint32_t with_a;            // 0: B only (the sparse solves), 1: also A right of a pivot row's own tile (the blocked solve)
};

// One warp per pivot row: mark the blocks its entries fall into.
__global__ void __launch_bounds__(256) bs_mark_kernel(BsArgs a) {
    const Layout& L = a.L;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t k = blockIdx.x * 8 + warp;
    if (k >= L.n_top) return;
    const uint32_t nr = L.n_rows, T = L.tile_cols, per_tile = T / a.cbw;
    const uint32_t kb = k / 128u, word = kb >> 5, bit = 1u << (kb & 31u);
    for (uint32_t t = a.with_a ? k / T + 1 : L.n_tiles_piv; t < L.n_tiles; ++t) {
        const uint32_t s = a.seg[(uint64_t)t * nr + k], e = a.seg[(uint64_t)t * nr + k + 1];
        const uint32_t cb0 = t < L.n_tiles_piv ? t * per_tile : a.n_cb_piv + (t - L.n_tiles_piv) * per_tile;
        for (uint32_t i = s + lane; i < e; i += 32) {
            const uint2 en = __ldg(a.ent + i);
            if (en.y == 0) continue;                       // padding
            uint32_t* w = a.bitmap + (uint64_t)(cb0 + (en.x & 0xffffu) / a.cbw) * a.kb_words + word;
            if (!(*reinterpret_cast<volatile uint32_t*>(w) & bit)) atomicOr(w, bit);
        }
    }
}

// One warp per column block: blocks below every bitmap word, count of the column block.
__global__ void __launch_bounds__(256) bs_rank_kernel(BsArgs a) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cb = blockIdx.x * 8 + warp;
    if (cb >= a.n_cb) return;
    uint32_t run = 0;
    for (uint32_t w0 = 0; w0 < a.kb_words; w0 += 32) {
        const uint32_t w = w0 + lane;
        const uint32_t c = w < a.kb_words ? __popc(a.bitmap[(uint64_t)cb * a.kb_words + w]) : 0u;
        uint32_t inc = c;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if ((int)lane >= o) inc += v; }
        if (w < a.kb_words) a.rank[(uint64_t)cb * a.kb_words + w] = run + inc - c;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) a.blk_base[cb] = run;
}

// One warp per column block (after the exclusive scan of blk_base): the k-blocks of its occupied blocks, ascending.
__global__ void __launch_bounds__(256) bs_list_kernel(BsArgs a) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cb = blockIdx.x * 8 + warp;
    if (cb >= a.n_cb) return;
    const uint32_t base = a.blk_base[cb];
    for (uint32_t w = lane; w < a.kb_words; w += 32) {
        uint32_t bits = a.bitmap[(uint64_t)cb * a.kb_words + w], pos = base + a.rank[(uint64_t)cb * a.kb_words + w];
        while (bits) { a.klist[pos++] = w * 32 + (__ffs(bits) - 1); bits &= bits - 1; }
    }
}


template <int NL>
__global__ void __launch_bounds__(256) operand_build_kernel(OpArgs a) {
    constexpr int CH = op_ch<NL>();
    extern __shared__ __align__(16) unsigned char op_sm[];
    const Layout& L = a.L;
    const uint32_t tid = threadIdx.x;
    const uint32_t T = L.tile_cols, nr = L.n_rows;
    const uint32_t kb = blockIdx.x, cc = blockIdx.y, z = blockIdx.z + a.z0;
    uint32_t tc, k0, orow0, okoff, pitch, rows;
    uint64_t plane;
    uint8_t* out = a.out;
    if (a.mode == OP_B) {
        tc = L.n_tiles_piv + z; k0 = kb * OP_KB; orow0 = z * T + cc * CH; okoff = k0; pitch = a.pitch; plane = a.plane; rows = a.rows;
    } else if (a.mode == OP_AT) {
        if (kb * OP_KB >= z * T) return;                            // only the pivot rows of earlier tiles reach tile z
        tc = z; k0 = kb * OP_KB; orow0 = cc * CH; okoff = k0; pitch = z * T; plane = (uint64_t)T * pitch; rows = T;
        if (!a.bitmap) out += a.at_off[z];
    } else {
        tc = z; k0 = z * T + kb * OP_KB; orow0 = z * T + cc * CH; okoff = kb * OP_KB; pitch = a.pitch; plane = a.plane; rows = 0xffffffffu;
    }
    if (orow0 >= rows) return;
    const uint32_t c_lo = cc * CH, c_hi = min(c_lo + CH, T), cw = c_hi - c_lo;     // the tile ends at T
    // block-sparse: the chunk's column blocks (CH / cbw of them); nothing to do if none of them is occupied for this k-block
    __shared__ uint32_t s_blk[8];
    if (a.bitmap) {
        if (tid < 8) {
            uint32_t id = 0xffffffffu;
            const uint32_t sub = tid;
            if (sub * a.cbw < cw) {
                const uint32_t cb = (a.mode == OP_B ? a.n_cb_piv + z * (T / a.cbw) : z * (T / a.cbw)) + c_lo / a.cbw + sub;
                const uint32_t w = a.bitmap[(uint64_t)cb * a.kb_words + (kb >> 5)], bit = kb & 31u;
                if ((w >> bit) & 1u) id = a.blk_base[cb] + a.rank[(uint64_t)cb * a.kb_words + (kb >> 5)] + __popc(w & ((1u << bit) - 1u)) - a.id0;
            }
            s_blk[sub] = id;
        }
        __syncthreads();
        bool any = false;
        for (uint32_t sub = 0; sub * a.cbw < cw; ++sub) any |= s_blk[sub] != 0xffffffffu;
        if (!any) return;
    }
    for (uint32_t i = tid * 16; i < NL * CH * OP_PITCH; i += 256 * 16) *reinterpret_cast<uint4*>(op_sm + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // in OP_ADT a chunk left of the block's rows holds nothing (A is upper triangular): it is written as zeros
    const bool empty = a.mode == OP_ADT && c_hi <= kb * OP_KB;
    if (!empty && k0 < L.n_top) {
        // the entries of the block's 128 pivot rows in this tile are ONE contiguous range (tile-major ent[]); an entry names
        // its row inside the block in bits 16-22. Two entries (16 bytes) per thread and pass: segments are padded to that.
        const uint32_t s = a.seg[(uint64_t)tc * nr + k0], e = a.seg[(uint64_t)tc * nr + min(k0 + OP_KB, L.n_top)];
        for (uint32_t q = s + tid * 2; q < e; q += 512) {
            const uint4 en2 = __ldg(reinterpret_cast<const uint4*>(a.ent + q));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t x = h ? en2.z : en2.x, v = h ? en2.w : en2.y;
                const uint32_t c = x & 0xffffu, i = (x >> 16) & 127u;
                if (v && c >= c_lo && c < c_hi) {
#pragma unroll
                    for (int l = 0; l < NL; ++l) op_sm[((size_t)l * CH + (c - c_lo)) * OP_PITCH + i] = (uint8_t)(v >> (8 * l));
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t w = tid; w < (uint32_t)(NL * CH * 8); w += 256) {
        const uint32_t part = w & 7u, c = (w >> 3) % CH, l = (w >> 3) / CH;
        if (c >= cw || orow0 + c >= rows) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(op_sm + ((size_t)l * CH + c) * OP_PITCH + part * 16);
        if (a.bitmap) {
            const uint32_t id = s_blk[c / a.cbw];
            if (id != 0xffffffffu) *reinterpret_cast<uint4*>(a.out + l * a.plane + ((uint64_t)id * a.cbw + c % a.cbw) * OP_KB + part * 16) = v;
        } else
        *reinterpret_cast<uint4*>(out + l * plane + (uint64_t)(orow0 + c) * pitch + okoff + part * 16) = v;
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


}  // namespace gcu
