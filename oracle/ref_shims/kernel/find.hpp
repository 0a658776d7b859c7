// TEST INFRASTRUCTURE (oracle/_ref build). Scalar stand-in for the reference's withheld SIMD reducer search
// (call site source/matrix.hpp:597-599; first-match semantics of the `#if 0` block, source/matrix.hpp:577-590):
// index of the first generator whose lead monomial divides the matrix monomial, or n.
#pragma once
#include <cstddef>
#include <cstdint>

namespace gamba
{
template <class MaskT, class ExpT, class IndT>
inline size_t find_multiplied_reducer_kernel(MaskT const mon_mask, ExpT const* mon_exp, size_t const exp_size,
                                             MaskT const* lm_masks, IndT const* lm_ind, ExpT const* basis_exps,
                                             size_t const n)
{
    for (size_t i = 0; i < n; ++i)
    {
        if (lm_masks[i] & ~mon_mask) continue;
        ExpT const* const e = basis_exps + static_cast<size_t>(lm_ind[i]) * exp_size;
        bool div = true;
        for (size_t k = 0; k < exp_size && div; ++k) div = e[k] <= mon_exp[k];
        if (div) return i;
    }
    return n;
}
}  // namespace gamba
