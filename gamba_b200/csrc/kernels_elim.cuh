// C|D-by-A|B elimination: every row to reduce is eliminated independently
// against the pivot rows (role of the reference's closed kernel/reduce_u16|u32,
// kernel/saxpy, kernel/modular_u16|u32 translation units, CMakeLists.txt:81-90).
//
// ONE CTA PER ROW IN FLIGHT (ELIM_CTA_WARPS warps sharing one accumulator tile),
// rows pulled from an atomic queue (most expensive first), as many CTAs resident
// per SM as shared memory allows. History:
//   v1  one CTA per row, 64-bit lanes, warp-per-segment        90 ms  (kat12-15 step 9)
//       -> 12 % active warps, 93 % L2 hits: latency-bound, not HBM-bound
//   v2  one warp per row, pipelined 32-entry chunks             35 ms
//       -> ~25 instructions per half-empty chunk: issue-bound per warp
//   v3  lane per segment + native 32-bit shared-memory atomics   54 ms
//   v4  one warp per row, chunk stream staged by cp.async, atomic lanes   51 ms
//       -> a single warp retires one 64-entry chunk per ~450 cycles whatever the
//          load depth: instruction latency per warp, so a row needs several warps
//   v5  (this file) v4's stream dealt round-robin to the warps of a CTA; the native
//       32-bit atomics make concurrent updates of the shared tile by all warps safe
//
// The row's accumulator is a tile in shared memory with DELAYED REDUCTION:
//   p < 2^16   one 32-bit lane per column; a multiply-add adds (m*v mod p) < 2^16
//              with ATOMS.ADD, the lane is reduced mod p only when it is read
//              (or after `norm_every` pivot rows, long before 2^32 is reached);
//   p < 2^31   a lo/hi pair of 32-bit lanes per column; (m*v mod p) is split into
//              its low 16 and high 15 bits, one ATOMS.ADD each (64-bit shared
//              atomics are a CAS loop on sm_100a, 32-bit ones are native and were
//              measured at 5.9 lanes/clk/SM vs 2.8 for LDS.64+STS.64,
//              tools/micro/atoms_bench.cu).
// Pivot-row entries are staged through a per-warp ring of shared-memory slots
// with cp.async (LDGSTS): ELIM_DEPTH x 512 B in flight per warp without holding
// registers (segments are sector-aligned and padded for it, layout.cuh). A
// lane-per-segment variant (every lane streaming its own pivot row with 32-byte
// loads, possible because the atomics make cross-row conflicts safe) was measured
// slower: 2 KB in flight per warp and uncoalesced sectors.
//
// The warp sweeps the column tiles left to right:
//   1. zero the tile, scatter the row's own entries of this tile;
//   2. DEFERRED pass: apply every pivot row found in earlier tiles (multiplier
//      list kept per warp in global memory) restricted to this tile;
//   3. pivot tiles only: walk the tile in WINDOWS of 32 pivot columns: read the
//      window, multiply by the precomputed INVERSE of the 32x32 diagonal block
//      (x = a (I+N)^-1, no dependent chain; only rows flagged in dnz are touched),
//      then apply the pivot rows with x != 0 to the rest of the tile;
//   4. non-pivot tiles: reduce the lanes mod p and write the row of D'.
//
// Three ways of running it (gamba_cuda.cu picks per step):
//   sparse path   all of the above for both halves of the row (operands of the products too big, option schur=0);
//   Schur mode    (a.mx) the non-pivot tiles only receive the row's own entries: D' = D + M B is a tensor-core
//                 product (kernels_schur_tc.cuh), the multipliers are recorded as u8 limb planes;
//   panel mode    (a.panel) one launch per column tile [t_lo, t_hi): step 2 is replaced by a tensor-core product
//                 cd = M[:, earlier tiles] * A[earlier tiles, tile] computed before the launch, so only step 3 is
//                 left - a chain of dependent windows per row, for which the row flags of the inverse diagonal
//                 block and the lane's column of its first flagged rows are fetched one window ahead.
#pragma once
#include "layout.cuh"

namespace gcu {

constexpr int ELIM_CTA_WARPS = 8;         // warps per row to reduce (template parameter WARPS: 8, or 4 when many rows
                                          // must be resident per SM - the walk over the windows is latency-bound; 32 when a
                                          // step has fewer rows than SMs against >= 65536 pivot rows: a row is ONE CTA, and
                                          // the rate of its deferred passes is warps x one chunk per ~450 cycles)
constexpr int ELIM_DEPTH = 4;             // cp.async ring slots per warp (512 B each)
constexpr int ELIM_SCRATCH_WORDS = ELIM_DEPTH * 128;   // per warp: the ring

struct ElimArgs {
    Layout L;
    Fp f;
    uint32_t mu32;               // floor(2^32 / p), Barrett constant for 32-bit operands (p < 2^16)
    uint32_t negp;               // 2^32 - p
    uint32_t norm_every;         // pivot rows a tile may absorb between normalisations
    const uint32_t* seg;
    const uint2* ent;
    const uint32_t* diag;
    const uint32_t* dnz;
    const uint32_t* first_piv;
    uint32_t* dprime;            // [n_local][ld_dprime] standard residues
    uint32_t* row_nz;            // [n_local] 1 iff the reduced row is non-zero
    uint32_t* list_k;            // [gridDim][n_top] pivot columns applied so far
    uint32_t* list_m;            // [gridDim][n_top] their multipliers (p - x)
    uint32_t* queue;             // atomic row counter
    uint32_t row_begin, row_step, n_local;  // rows r = row_begin + i * row_step, i < n_local
    const uint64_t* row_ptr;     // block-form CSR row pointers (for the fma_top invariant)
    unsigned long long* stats;   // [0] pivot rows applied, [1] multiply-adds done, [2] fma_top
    // Schur mode (kernels_schur.cuh): the non-pivot tiles are NOT eliminated here; the multipliers are
    // recorded as u8 limb planes mx[limb][local row][k] and D' = D + M B runs on the tensor cores
    uint8_t* mx;                 // NULL = eliminate the non-pivot tiles here (sparse path)
    uint64_t mx_plane;
    uint32_t mx_kp, mx_nl;
    // Panel mode (kernels_schur_tc.cuh as a blocked triangular solve): one launch walks the windows of the column
    // tiles [t_lo, t_hi) only; what the pivots of EARLIER tiles add to tile t_lo was computed on the tensor cores,
    // cd[local row][tile-relative column] = sum_k M[row][k] A[k][col], so there is no deferred pass and no list
    uint32_t t_lo, t_hi;         // 0, n_tiles outside panel mode
    const uint32_t* cd;          // NULL outside panel mode (and for the first pivot tile)
    uint32_t cd_ld, panel;
};

// ---- accumulator policies -------------------------------------------------------------------

template <bool SMALLP> struct Acc;

__device__ __forceinline__ void red_shared_add(uint32_t saddr, uint32_t val) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(val) : "memory");
}

// the same, skipped when val == 0: a predicated instruction, no branch (the blocked kernels issue 16 of these per lane and chunk)
__device__ __forceinline__ void red_shared_add_nz(uint32_t saddr, uint32_t val, uint32_t nz) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" ::"r"(saddr), "r"(val), "r"(nz) : "memory");
}

template <> struct Acc<true> {   // p < 2^16: one u32 lane per column
    static constexpr int WORDS = 1;
    // m*v mod p up to one multiple of p: the result is < 2p (the lane is reduced when it is read)
    static __device__ __forceinline__ uint32_t mulred(uint32_t m, uint32_t v, const ElimArgs& a) {
        const uint32_t x = m * v;                       // < 2^32
        return x + __umulhi(x, a.mu32) * a.negp;        // x - q*p, one IMAD
    }
    static __device__ __forceinline__ void add(uint32_t acc_sa, uint32_t c, uint32_t val) { red_shared_add(acc_sa + c * 4, val); }
    static __device__ __forceinline__ void add_nz(uint32_t acc_sa, uint32_t val) { red_shared_add_nz(acc_sa, val, val); }
    static __device__ __forceinline__ void set(uint32_t* acc, uint32_t c, uint32_t val) { acc[c] = val; }
    static __device__ __forceinline__ uint32_t get(const uint32_t* acc, uint32_t c, const ElimArgs& a) {
        const uint32_t x = acc[c];
        uint32_t r = x - __umulhi(x, a.mu32) * a.f.p;
        return r >= a.f.p ? r - a.f.p : r;
    }
};

template <> struct Acc<false> {  // p < 2^31: lo/hi pair of u32 lanes per column
    static constexpr int WORDS = 2;
    static __device__ __forceinline__ uint32_t mulred(uint32_t m, uint32_t v, const ElimArgs& a) {
        return fp_reduce64((uint64_t)m * v, a.f);
    }
    static __device__ __forceinline__ void add(uint32_t acc_sa, uint32_t c, uint32_t val) {
        red_shared_add(acc_sa + c * 8, val & 0xffffu);
        red_shared_add(acc_sa + c * 8 + 4, val >> 16);
    }
    static __device__ __forceinline__ void add_nz(uint32_t acc_sa, uint32_t val) {
        red_shared_add_nz(acc_sa, val & 0xffffu, val);
        red_shared_add_nz(acc_sa + 4, val >> 16, val);
    }
    static __device__ __forceinline__ void set(uint32_t* acc, uint32_t c, uint32_t val) {
        acc[2 * c] = val & 0xffffu; acc[2 * c + 1] = val >> 16;
    }
    static __device__ __forceinline__ uint32_t get(const uint32_t* acc, uint32_t c, const ElimArgs& a) {
        const uint2 w = *reinterpret_cast<const uint2*>(acc + 2 * c);
        return fp_reduce64((uint64_t)w.x + ((uint64_t)w.y << 16), a.f);
    }
};

// Bring every lane of the tile back below p (delayed reduction: done rarely).
template <bool SMALLP, int ELIM_CTA_THREADS>
__device__ __forceinline__ void elim_normalise(uint32_t* acc, uint32_t tw, const ElimArgs& a) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tw; i += ELIM_CTA_THREADS) Acc<SMALLP>::set(acc, i, Acc<SMALLP>::get(acc, i, a));
    __syncthreads();
}

// ---- asynchronous staging of pivot-row entries (LDGSTS) ---------------------------------------

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gmem, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
        "@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}\n" ::"r"(smem_addr), "l"(gmem), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Apply the pivot-row segments selected by `mask` (lane i holds segment i: entries [s, s+len),
// multiplier m). The segments are walked in lane order and cut into CHUNKS of 64 consecutive
// entries (512 B, 16 B per lane) that run through the warp's ring of ELIM_DEPTH shared-memory
// slots filled by cp.async: ELIM_DEPTH x 512 B in flight per warp at no register cost. A chunk is
// all lanes on consecutive entries of ONE pivot row: coalesced, and each lane reads back exactly
// the 16 B it copied itself. Two warp-uniform cursors (issue side, consume side) walk the same
// chunk sequence, so the per-chunk bookkeeping is a compare and an add (profiles/r01_v5: the
// previous form recomputed the chunk owner with ballot/popc/shuffles and was issue-bound).
struct ChunkCursor {
    uint32_t mask, s, len, m, off;
};

__device__ __forceinline__ bool cursor_next(ChunkCursor& c, uint32_t s, uint32_t len, uint32_t m) {
    if (c.off >= c.len) {
        if (c.mask == 0) return false;
        const int o = __ffs(c.mask) - 1;
        c.mask &= c.mask - 1;
        c.s = __shfl_sync(0xffffffffu, s, o);
        c.len = __shfl_sync(0xffffffffu, len, o);
        c.m = __shfl_sync(0xffffffffu, m, o);
        c.off = 0;
    }
    return true;
}

template <bool SMALLP>
__device__ __forceinline__ void elim_apply_segments(uint32_t acc_sa, uint32_t ring_sa, const ElimArgs& a, uint32_t s,
                                                    uint32_t len, uint32_t m, uint32_t mask, uint32_t lane) {
    if (mask == 0) return;
    static_assert((ELIM_DEPTH & (ELIM_DEPTH - 1)) == 0, "ring depth must be a power of two");
    const uint32_t ring_lane = ring_sa + lane * 16;
    ChunkCursor ci{mask, 0, 0, 0, 0}, cc{mask, 0, 0, 0, 0};
    uint32_t in = 0, cn = 0;
    auto issue = [&]() {
        if (cursor_next(ci, s, len, m)) {
            const uint32_t off = ci.off + lane * 2;
            cp_async16(ring_lane + (in & (ELIM_DEPTH - 1)) * 512, a.ent + ci.s + off, off < ci.len);
            ci.off += 64;
        }
        cp_async_commit();
        ++in;
    };
#pragma unroll
    for (int i = 0; i + 1 < ELIM_DEPTH; ++i) issue();
    while (cursor_next(cc, s, len, m)) {
        issue();
        cp_async_wait<ELIM_DEPTH - 1>();
        const uint32_t off = cc.off + lane * 2;
        if (off < cc.len) {   // len is a multiple of 4, off is even: both entries are inside
            uint4 q;          // padding entries carry value 0 and the sink column: adding 0 there is harmless
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                         : "r"(ring_lane + (cn & (ELIM_DEPTH - 1)) * 512) : "memory");
            Acc<SMALLP>::add(acc_sa, q.x & 0xffffu, Acc<SMALLP>::mulred(cc.m, q.y, a));   // bits 16-20: window slot of the pivot row
            Acc<SMALLP>::add(acc_sa, q.z & 0xffffu, Acc<SMALLP>::mulred(cc.m, q.w, a));
        }
        cc.off += 64;
        ++cn;
    }
    cp_async_wait<0>();
}

template <bool SMALLP, int ELIM_CTA_WARPS>
__global__ void __launch_bounds__(ELIM_CTA_WARPS * 32) elim_kernel(ElimArgs a) {
    constexpr int ELIM_CTA_THREADS = ELIM_CTA_WARPS * 32;
    extern __shared__ __align__(16) unsigned char elim_smem[];
    __shared__ uint32_t s_q;
    const Layout& L = a.L;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int W = Acc<SMALLP>::WORDS;
    const uint32_t acc_words = ((L.tile_cols + 1) * W + 3u) & ~3u;               // +1: the sink lane
    uint32_t* acc = reinterpret_cast<uint32_t*>(elim_smem);
    uint32_t* ring = acc + acc_words + warp * ELIM_SCRATCH_WORDS;
    const uint32_t acc_sa = (uint32_t)__cvta_generic_to_shared(acc);
    const uint32_t ring_sa = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t p = a.f.p, nr = L.n_rows;
    const uint64_t p2 = a.f.p2;
    uint32_t* lk = a.list_k + (uint64_t)blockIdx.x * L.n_top;
    uint32_t* lm = a.list_m + (uint64_t)blockIdx.x * L.n_top;
    unsigned long long fma = 0, npiv = 0, fma_top = 0, fma_top_pend = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_q = atomicAdd(a.queue, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= a.n_local) break;
        const uint32_t li = a.n_local - 1 - q;                 // later rows have smaller leads: more work
        const uint32_t r = a.row_begin + li * a.row_step;      // index among the rows to reduce
        const uint32_t grow = L.n_top + r;
        const uint32_t first = a.first_piv[r];
        const uint32_t t_first = first < L.n_top ? first / L.tile_cols : L.n_tiles_piv;
        uint32_t nlist = 0;
        bool nz = false;

        for (uint32_t t = max(t_first, a.t_lo); t < a.t_hi; ++t) {
            const uint32_t t0 = lay_tile_start(L, t), tw = lay_tile_width(L, t);
            const uint32_t* seg_s = a.seg + (uint64_t)t * nr;
            const uint32_t* seg_e = seg_s + 1;          // tile-major table: a segment ends where the next row's starts
            // the previous tile is read (last window / write-out of D') by other threads than those that
            // zero the same words here (two words per column for p >= 2^16): cyclic9-31 lost updates without this
            __syncthreads();
            for (uint32_t i = tid; i < acc_words; i += ELIM_CTA_THREADS) acc[i] = 0;
            __syncthreads();
            uint32_t absorbed = 1;   // rows added into this tile since the last normalisation
            if (a.cd && t < L.n_tiles_piv) {   // panel mode: the earlier pivots' contribution, then the row's own entries on top
                const uint32_t* cdr = a.cd + (uint64_t)li * a.cd_ld;
                for (uint32_t i = tid; i < tw; i += ELIM_CTA_THREADS) Acc<SMALLP>::set(acc, i, __ldcg(cdr + i));
                __syncthreads();
                const uint32_t s = seg_s[grow], e = seg_e[grow];
                for (uint32_t k = s + tid; k < e; k += ELIM_CTA_THREADS) { const uint2 en = __ldg(a.ent + k); Acc<SMALLP>::add(acc_sa, en.x & 0xffffu, en.y); }
                absorbed = 2;
            } else {   // the row itself
                const uint32_t s = seg_s[grow], e = seg_e[grow];
                for (uint32_t k = s + tid; k < e; k += ELIM_CTA_THREADS) { const uint2 en = __ldg(a.ent + k); Acc<SMALLP>::set(acc, en.x & 0xffffu, en.y); }
            }
            __syncthreads();
            // deferred pass over the pivots found in earlier tiles
            // (batches of 32 pivots are dealt to the warps; they commute, no barrier between them)
            const uint32_t ndefer = ((a.mx && t >= L.n_tiles_piv) || a.panel) ? 0u : nlist;     // Schur mode: B half on the tensor cores
            // the list entries and segment bounds of a batch are fetched while the batch before it streams (two dependent loads:
            // ~2.5 us per batch of 256 pivots otherwise - 150 of 167 ms on a kat15-15 step of 13 rows against 410 k pivots)
            auto fetch = [&](uint32_t base0, uint32_t& s, uint32_t& len, uint32_t& m) {
                const uint32_t i = base0 + warp * 32 + lane;
                s = 0; len = 0; m = 0;
                if (i < ndefer) { const uint32_t k = lk[i]; m = lm[i]; s = seg_s[k]; len = seg_e[k] - s; }
            };
            uint32_t s_n, len_n, m_n;
            fetch(0, s_n, len_n, m_n);
            for (uint32_t base0 = 0; base0 < ndefer; base0 += 32 * ELIM_CTA_WARPS) {
                if (absorbed + 32 * ELIM_CTA_WARPS > a.norm_every) { elim_normalise<SMALLP, ELIM_CTA_THREADS>(acc, tw, a); absorbed = 1; }
                absorbed += 32 * ELIM_CTA_WARPS;
                const uint32_t s = s_n, len = len_n, m = m_n;
                fetch(base0 + 32 * ELIM_CTA_WARPS, s_n, len_n, m_n);
                fma += len;
                elim_apply_segments<SMALLP>(acc_sa, ring_sa, a, s, len, m, __ballot_sync(0xffffffffu, len != 0), lane);
            }
            __syncthreads();

            if (t < L.n_tiles_piv) {
                const uint32_t nwin = (tw + WINDOW - 1) / WINDOW;
                uint32_t g = (t == t_first ? (first - t0) / WINDOW : 0);
                // segment pointers of the window's 32 pivot rows, fetched one window ahead; so are the row flags of the
                // window's inverse diagonal block and this lane's column of its first 4 flagged rows (the walk is a chain
                // of dependent steps: every global load on it costs ~1 us per window)
                uint32_t ps = 0, pe = 0;
                if (g < nwin && t0 + g * WINDOW + lane < L.n_top) { ps = seg_s[t0 + g * WINDOW + lane]; pe = seg_e[t0 + g * WINDOW + lane]; }
                uint32_t dzA = 0, dzB = 0, pvA0 = 0, pvA1 = 0, pvA2 = 0, pvA3 = 0;
                auto fetch_diag = [&](uint32_t win, uint32_t dz, uint32_t& v0, uint32_t& v1, uint32_t& v2, uint32_t& v3) {
                    const uint32_t* dg = a.diag + (uint64_t)win * WINDOW * WINDOW + lane;
                    uint32_t rest = dz;
                    v0 = v1 = v2 = v3 = 0;
                    if (rest) { v0 = __ldg(dg + (__ffs(rest) - 1) * WINDOW); rest &= rest - 1; }
                    if (rest) { v1 = __ldg(dg + (__ffs(rest) - 1) * WINDOW); rest &= rest - 1; }
                    if (rest) { v2 = __ldg(dg + (__ffs(rest) - 1) * WINDOW); rest &= rest - 1; }
                    if (rest) { v3 = __ldg(dg + (__ffs(rest) - 1) * WINDOW); }
                };
                if (g < nwin) {
                    const uint32_t win = (t0 + g * WINDOW) / WINDOW;
                    dzA = a.dnz[win];
                    if (g + 1 < nwin) dzB = a.dnz[win + 1];
                    fetch_diag(win, dzA, pvA0, pvA1, pvA2, pvA3);
                }
                for (; g < nwin; ++g) {
                    // every warp derives the window's multipliers itself (identical results, no broadcast)
                    const uint32_t k0 = t0 + g * WINDOW;
                    const uint32_t c = g * WINDOW + lane;
                    const uint32_t s = ps, e = pe;
                    if (g + 1 < nwin && k0 + WINDOW + lane < L.n_top) { ps = seg_s[k0 + WINDOW + lane]; pe = seg_e[k0 + WINDOW + lane]; }
                    const uint32_t dz = dzA, pv0 = pvA0, pv1 = pvA1, pv2 = pvA2, pv3 = pvA3;
                    dzA = dzB;
                    if (g + 1 < nwin) fetch_diag(k0 / WINDOW + 1, dzA, pvA0, pvA1, pvA2, pvA3);
                    dzB = g + 2 < nwin ? a.dnz[k0 / WINDOW + 2] : 0u;
                    fma_top += fma_top_pend; fma_top_pend = 0;
                    const uint32_t av = c < tw ? Acc<SMALLP>::get(acc, c, a) : 0u;
                    const uint32_t nza = __ballot_sync(0xffffffffu, av != 0);
                    if (nza == 0) continue;           // CTA-uniform: all warps read the same lanes
                    uint32_t x = av;
                    const uint32_t need = nza & dz;
                    if (need) {   // x = a * (I + M): add a_i * M[i][lane] for the flagged rows i
                        const uint32_t* dg = a.diag + (uint64_t)(k0 / WINDOW) * WINDOW * WINDOW + lane;
                        // four independent partial sums: the chain of dependent 64-bit multiply-adds is what a window waits for
                        uint64_t xa0 = av, xa1 = 0, xa2 = 0, xa3 = 0;
                        uint32_t rest = dz;
                        int i;
                        if (rest) { i = __ffs(rest) - 1; rest &= rest - 1; if ((need >> i) & 1u) xa0 = fp_mac<SMALLP>(xa0, __shfl_sync(0xffffffffu, av, i), pv0, p2); }
                        if (rest) { i = __ffs(rest) - 1; rest &= rest - 1; if ((need >> i) & 1u) xa1 = fp_mac<SMALLP>(xa1, __shfl_sync(0xffffffffu, av, i), pv1, p2); }
                        if (rest) { i = __ffs(rest) - 1; rest &= rest - 1; if ((need >> i) & 1u) xa2 = fp_mac<SMALLP>(xa2, __shfl_sync(0xffffffffu, av, i), pv2, p2); }
                        if (rest) { i = __ffs(rest) - 1; rest &= rest - 1; if ((need >> i) & 1u) xa3 = fp_mac<SMALLP>(xa3, __shfl_sync(0xffffffffu, av, i), pv3, p2); }
                        rest &= need;
                        while (rest) {                // more than 4 flagged rows: the others on demand, 4 independent loads at a time
                            int ii[4]; uint32_t mm[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                ii[u] = rest ? __ffs(rest) - 1 : -1;
                                if (rest) rest &= rest - 1;
                                mm[u] = ii[u] >= 0 ? __ldg(dg + ii[u] * WINDOW) : 0u;
                            }
                            if (ii[0] >= 0) xa0 = fp_mac<SMALLP>(xa0, __shfl_sync(0xffffffffu, av, ii[0]), mm[0], p2);   // p < 2^16: 32 products fit 64 bits
                            if (ii[1] >= 0) xa1 = fp_mac<SMALLP>(xa1, __shfl_sync(0xffffffffu, av, ii[1]), mm[1], p2);
                            if (ii[2] >= 0) xa2 = fp_mac<SMALLP>(xa2, __shfl_sync(0xffffffffu, av, ii[2]), mm[2], p2);
                            if (ii[3] >= 0) xa3 = fp_mac<SMALLP>(xa3, __shfl_sync(0xffffffffu, av, ii[3]), mm[3], p2);
                        }
                        x = fp_reduce64(xa0 + xa1 + xa2 + xa3, a.f);   // each partial sum < p^2 < 2^62
                    }
                    const uint32_t nzm = __ballot_sync(0xffffffffu, x != 0);
                    if (nzm == 0) continue;           // CTA-uniform
                    uint32_t m = 0;
                    if (x) {
                        m = p - x;
                        if (warp == 0) {
                            const uint32_t i = nlist + __popc(nzm & lt);
                            if (!a.panel) { lk[i] = k0 + lane; lm[i] = m; }
                            if (a.mx) {
                                uint8_t* mp = a.mx + (uint64_t)li * a.mx_kp + k0 + lane;
                                for (uint32_t l = 0; l < a.mx_nl; ++l) mp[l * a.mx_plane] = (uint8_t)(m >> (8 * l));
                            }
                            fma_top_pend = a.row_ptr[k0 + lane + 1] - a.row_ptr[k0 + lane];   // consumed one window later: off the chain
                        }
                    }
                    nlist += __popc(nzm);
                    if (absorbed + 32 > a.norm_every) { elim_normalise<SMALLP, ELIM_CTA_THREADS>(acc, tw, a); absorbed = 1; }
                    absorbed += 32;
                    // the window's non-zero pivot rows are dealt to the warps by rank
                    const uint32_t len = x ? e - s : 0u;
                    const bool mine = len != 0 && (__popc(nzm & lt) % ELIM_CTA_WARPS) == warp;
                    if (mine) fma += len;
                    elim_apply_segments<SMALLP>(acc_sa, ring_sa, a, s, len, m, __ballot_sync(0xffffffffu, mine), lane);
                    __syncthreads();                  // the next window is read after all warps' updates
                }
            } else {
                uint32_t* out = a.dprime + (uint64_t)li * L.ld_dprime + (t0 - L.n_top);
                for (uint32_t i = tid; i < tw; i += ELIM_CTA_THREADS) {
                    const uint32_t v = Acc<SMALLP>::get(acc, i, a);
                    out[i] = v;
                    nz |= v != 0;
                }
            }
        }
        const int any = __syncthreads_or(nz ? 1 : 0);
        if (tid == 0 && !a.mx) a.row_nz[li] = any ? 1u : 0u;      // Schur mode: the product kernel sets the flag
        npiv += tid == 0 ? nlist : 0;
    }
    for (int o = 16; o; o >>= 1) fma += __shfl_xor_sync(0xffffffffu, fma, o);
    if (lane == 0 && fma) atomicAdd(a.stats + 1, fma);
    if (npiv) atomicAdd(a.stats + 0, npiv);
    fma_top += fma_top_pend;
    if (warp == 0 && fma_top) atomicAdd(a.stats + 2, fma_top);
}

}  // namespace gcu
