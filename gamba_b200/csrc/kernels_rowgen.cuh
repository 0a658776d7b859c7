// Row generation of symbolic preprocessing on the device (role of the reference's
// matrix_f4::create_multiplied_poly_matrix_row + its monomial hash table, source/matrix.hpp:387-444, called from
// insert_spairs :215-347 and symbolic_preprocessing :469-566).
//
// A row of the step matrix is a basis polynomial times a monomial: every term's exponent vector is added to the
// multiplier's, the product is looked up in the step's monomial table (inserted if new) and the term becomes the
// monomial's 32-bit handle. On the host this is one dependent chain of cache misses per term (slot -> hash -> exponents)
// into a table of up to 5e5 monomials - 25-30 s of a kat15-15 run on 16 threads, and the 4 bytes per term then cross PCIe
// (53 GB over that run). On the device the table (<= 40 MB) lives in L2, a term is one thread, and the handles are
// written straight into the array the staging kernels read (GAMBA_CUDA_STEP_DEVICE_COLIDX): no O(nnz) host work or
// traffic is left; what goes back to the host per batch is the new monomials (for its reducer search) and one lead
// handle per row.
//
//   basis table  bex[monomial][sd] u16, bhv[monomial] u32: the driver's basis monomials, appended as its table grows
//   term store   pmon[term] u32: the basis polynomials as handles into that table, uploaded once per polynomial (4 B / term)
//   step table   slots[cap_slots] (handle + 1 | 0 = empty | PENDING), hv[handle], ex[handle][sd]; open addressing,
//                linear probing, additive hashes (hash(a * b) = hash(a) + hash(b), as source/monomial.hpp:80-90,126-131)
//   sd           exponent-vector stride rounded up to an even number of u16 (32-bit compares)
#pragma once
#include <cstdint>

namespace gcu {

constexpr uint32_t RG_PENDING = 0xffffffffu;
constexpr int RG_THREADS = 256;
constexpr int RG_MAX_SD = 128;            // u16 per exponent vector the kernels support (2 degree slots + 126 variables)

struct RowGenArgs {
    uint32_t sd;                          // u16 per exponent vector (even)
    const uint16_t* bex;                  // [basis monomial][sd] the driver's basis monomials
    const uint32_t* bhv;                  // [basis monomial] their hashes
    const uint32_t* pmon;                 // term store: basis-monomial handle of every term of the uploaded polynomials
    uint32_t n_rows;
    const uint64_t* first_term;           // [n_rows] first term of the row's polynomial in the term store
    const uint64_t* row_off;              // [n_rows + 1] where the row's handles go in row_mon
    const uint16_t* q;                    // [n_rows][sd] multipliers
    const uint32_t* hq;                   // [n_rows]
    uint32_t* slots;
    uint32_t mask;
    uint32_t* hv;
    uint16_t* ex;
    uint32_t* count;                      // monomials in the table
    uint32_t cap;                         // handles the arrays hold
    uint32_t* overflow;
    uint32_t* row_mon;
    uint32_t* lead;                       // [n_rows] handle of the row's first (leading) term
};

// One CTA walks rows blockIdx.x, blockIdx.x + gridDim.x, ...; one thread per term. W = 32-bit words of an exponent vector the
// thread keeps in registers (>= sd / 2).
template <int W>
__global__ void __launch_bounds__(RG_THREADS) rowgen_kernel(RowGenArgs a) {
    __shared__ uint32_t s_q[RG_MAX_SD / 2];
    const uint32_t w = a.sd / 2;                      // 32-bit words per exponent vector
    for (uint32_t row = blockIdx.x; row < a.n_rows; row += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < w) s_q[threadIdx.x] = reinterpret_cast<const uint32_t*>(a.q + (uint64_t)row * a.sd)[threadIdx.x];
        __syncthreads();
        const uint64_t t0 = a.first_term[row], o0 = a.row_off[row];
        const uint32_t len = (uint32_t)(a.row_off[row + 1] - o0), hq = a.hq[row];
        for (uint32_t k = threadIdx.x; k < len; k += RG_THREADS) {
            uint32_t e[W];
            const uint32_t bm = a.pmon[t0 + k];
            const uint32_t* src = reinterpret_cast<const uint32_t*>(a.bex + (uint64_t)bm * a.sd);
            // two 16-bit exponents per word: no carry between the halves (exponents stay below 2^15, source/io.cpp:208-214)
#pragma unroll
            for (int j = 0; j < W; ++j) e[j] = (uint32_t)j < w ? src[j] + s_q[j] : 0u;
            const uint32_t h = a.bhv[bm] + hq;
            uint32_t s = h & a.mask, id;
            for (;;) {
                uint32_t v = *reinterpret_cast<volatile uint32_t*>(a.slots + s);
                if (v == 0) {
                    v = atomicCAS(a.slots + s, 0u, RG_PENDING);
                    if (v == 0) {                     // ours: take the next handle, fill it, publish it
                        id = atomicAdd(a.count, 1u);
                        if (id >= a.cap) { *a.overflow = 1u; id = 0; atomicExch(a.slots + s, 0u); break; }
                        a.hv[id] = h;
                        uint32_t* dst = reinterpret_cast<uint32_t*>(a.ex + (uint64_t)id * a.sd);
#pragma unroll
                        for (int j = 0; j < W; ++j) if ((uint32_t)j < w) dst[j] = e[j];
                        __threadfence();
                        atomicExch(a.slots + s, id + 1u);
                        break;
                    }
                }
                if (v == RG_PENDING) continue;        // being written by another thread: look again
                const uint32_t cand = v - 1u;
                bool same = __ldcg(a.hv + cand) == h;
                if (same) {
                    const uint32_t* ce = reinterpret_cast<const uint32_t*>(a.ex + (uint64_t)cand * a.sd);
#pragma unroll
                    for (int j = 0; j < W; ++j) if ((uint32_t)j < w && __ldcg(ce + j) != e[j]) same = false;
                }
                if (same) { id = cand; break; }
                s = (s + 1u) & a.mask;
            }
            a.row_mon[o0 + k] = id;
            if (k == 0) a.lead[row] = id;
        }
    }
}

// Rebuild the slots after the table grew: handles 0 .. n-1 are distinct monomials.
__global__ void __launch_bounds__(256) rowgen_rehash_kernel(uint32_t* slots, uint32_t mask, const uint32_t* hv, uint32_t n) {
    const uint32_t id = blockIdx.x * 256 + threadIdx.x;
    if (id >= n) return;
    uint32_t s = hv[id] & mask;
    while (atomicCAS(slots + s, 0u, id + 1u) != 0u) s = (s + 1u) & mask;
}

}  // namespace gcu
