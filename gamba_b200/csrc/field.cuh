// F_p arithmetic on the device, p an odd prime < 2^31.
//
// Number representation: standard residues in [0,p) everywhere on the device; the
// reference's Montgomery form (source/field.hpp:41-51, source/montgomery.hpp:39-65)
// is converted at the boundary (kernels_prep.cuh / kernels_echelon.cuh) when the
// engine is created with GAMBA_CUDA_REPR_MONTGOMERY.
//
// Sparse-row accumulators are 64-bit lanes with delayed reduction: a lane holds a
// value < p^2 and a multiply-add adds m*v < p^2, then subtracts p^2 once if needed
// (sum < 2 p^2 < 2^63). For p < 2^16 products are < 2^32 and a 64-bit lane takes
// 2^32 of them, so the correction is compiled out (the reference's max_fma,
// source/field.hpp:81-83, plays the same role for its AVX2 lanes).
#pragma once
#include <cstdint>

namespace gcu {

struct Fp {
    uint32_t p;
    uint32_t small;     // 1 iff p < 2^16
    uint64_t p2;        // p*p
    uint64_t mu;        // floor((2^64 - 1) / p), Barrett constant for 64-bit operands
    uint32_t mont_bits; // 0 (standard), 32 or 64: R = 2^mont_bits at the ABI
    uint32_t r_mod_p;   // R mod p      (x_std -> x_mont: x * R)
    uint32_t rinv_mod_p;// R^-1 mod p   (x_mont -> x_std: x * R^-1)
};

__host__ __device__ __forceinline__ uint64_t fp_mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// x mod p for any 64-bit x (Barrett, one correction step covers the floor error)
__host__ __device__ __forceinline__ uint32_t fp_reduce64(uint64_t x, const Fp& f) {
    uint64_t q = fp_mulhi64(x, f.mu);
    uint64_t r = x - q * f.p;
    if (r >= f.p) r -= f.p;
    if (r >= f.p) r -= f.p;
    return (uint32_t)r;
}

__host__ __device__ __forceinline__ uint32_t fp_mul(uint32_t a, uint32_t b, const Fp& f) {
    return fp_reduce64((uint64_t)a * b, f);
}

// acc + m*v with the delayed-reduction invariant acc < p^2 (or unconstrained when SMALLP)
template <bool SMALLP>
__device__ __forceinline__ uint64_t fp_mac(uint64_t acc, uint32_t m, uint32_t v, uint64_t p2) {
    uint64_t t = acc + (uint64_t)m * v;
    if (!SMALLP) { if (t >= p2) t -= p2; }
    return t;
}

__host__ __device__ inline uint32_t fp_inv(uint32_t a, uint32_t p) {
    // extended Euclid on standard representatives (as source/field.hpp:90-121)
    int64_t t = 0, nt = 1, r = p, nr = a % p;
    while (nr) {
        int64_t q = r / nr, x = t - q * nt; t = nt; nt = x;
        x = r - q * nr; r = nr; nr = x;
    }
    return (uint32_t)(t < 0 ? t + p : t);
}

// 32-bit variant for device code (64-bit division is emulated and slow on the GPU)
__host__ __device__ inline uint32_t fp_inv32(uint32_t a, uint32_t p) {
    int32_t t = 0, nt = 1;
    uint32_t r = p, nr = a % p;
    while (nr) {
        const uint32_t q = r / nr;
        const int32_t x = t - (int32_t)q * nt; t = nt; nt = x;
        const uint32_t y = r - q * nr; r = nr; nr = y;
    }
    return t < 0 ? (uint32_t)(t + (int32_t)p) : (uint32_t)t;
}

// a^(p-2) mod p: branch-light chain of Barrett multiplications (the extended Euclid above needs
// ~45 dependent integer divisions, several microseconds on one GPU thread)
__host__ __device__ inline uint32_t fp_inv_fermat(uint32_t a, const Fp& f) {
    uint32_t r = 1, b = a;
    for (uint32_t e = f.p - 2; e; e >>= 1) {
        if (e & 1) r = fp_mul(r, b, f);
        b = fp_mul(b, b, f);
    }
    return r;
}

inline Fp fp_make(uint32_t p, uint32_t mont_bits) {
    Fp f{};
    f.p = p; f.small = p < 65536u; f.p2 = (uint64_t)p * p;
    f.mu = ~0ull / p;
    f.mont_bits = mont_bits;
    uint64_t r = 1;
    for (uint32_t i = 0; i < mont_bits; ++i) r = (r * 2) % p;
    f.r_mod_p = (uint32_t)r;
    f.rinv_mod_p = fp_inv((uint32_t)r, p);
    return f;
}

}  // namespace gcu
