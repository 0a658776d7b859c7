// Reference-side binding #2: `source/linalg.hpp` (final inter-reduction, call site source/reduce.hpp:51-52) on top of
// gamba-b200's C ABI. See linalg_v3.hpp in this directory. linalg<Matrix> is not a friend of matrix_f4
// (source/matrix.hpp:173-177): it goes through the public accessors (source/matrix.hpp:109-144); the rows were already
// rewritten in place to uint32 column indices (source/matrix.hpp:702-719).
#pragma once
#include <chrono>
#include <cstdint>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "gamba_cuda.h"
#include "params.hpp"
#include "stats.hpp"

namespace gamba
{

template <class MatrixType>
class linalg
{
public:
    using matrix_type = MatrixType;
    using coeff_type  = typename matrix_type::coeff_type;
    using index_type  = typename matrix_type::index_type;
    using basis_type  = typename matrix_type::basis_type;

    explicit linalg(matrix_type& matrix)
    {
        gamba_cuda_config cfg{};
        cfg.p           = static_cast<uint32_t>(matrix.m_field.n);
        cfg.coeff_bytes = sizeof(coeff_type);
        cfg.repr        = GAMBA_CUDA_REPR_MONTGOMERY;
        cfg.seed        = 0;
        cfg.world_size  = 1;
        if (gamba_cuda_create(&cfg, &m_handle) != 0)
            throw std::runtime_error(std::string("gamba_cuda_create: ") + gamba_cuda_last_error());
    }

    linalg(linalg const&)            = delete;
    linalg& operator=(linalg const&) = delete;

    ~linalg() { gamba_cuda_destroy(m_handle); }

    void reduce(matrix_type& matrix, gamba_params const& /*params*/)
    {
        size_t const num_cols = matrix.num_cols();
        size_t const num_top  = matrix.num_top_rows();
        size_t const num_bot  = matrix.num_bottom_rows();

        auto indices = [](auto const* mons) { return reinterpret_cast<index_type const*>(mons); };

        std::vector<uint8_t> is_pivot(num_cols, 0);
        for (size_t i = 0; i < num_top; ++i)
            is_pivot[indices(matrix.top_monomials(i))[0]] = 1;
        std::vector<uint32_t> block_col(num_cols), nonpiv_col;
        uint32_t next_piv = 0, next_nonpiv = static_cast<uint32_t>(num_top);
        for (size_t c = 0; c < num_cols; ++c)
        {
            block_col[c] = is_pivot[c] ? next_piv++ : next_nonpiv++;
            if (!is_pivot[c])
                nonpiv_col.push_back(static_cast<uint32_t>(c));
        }

        std::vector<uint64_t> row_ptr{0};
        std::vector<uint32_t> col, coeff_id, cf_len;
        std::vector<coeff_type const*> cf_ptr;
        std::unordered_map<coeff_type const*, uint32_t> ids;
        auto emit = [&](index_type const* idx, size_t const len, coeff_type const* const cfs) {
            for (size_t k = 0; k < len; ++k)
                col.push_back(block_col[idx[k]]);
            row_ptr.push_back(col.size());
            auto const [it, fresh] = ids.try_emplace(cfs, static_cast<uint32_t>(cf_ptr.size()));
            if (fresh)
            {
                cf_ptr.push_back(cfs);
                cf_len.push_back(static_cast<uint32_t>(len));
            }
            coeff_id.push_back(it->second);
        };
        for (size_t i = 0; i < num_top; ++i)
            emit(indices(matrix.top_monomials(i)), matrix.top_size(i), matrix.top_coeffs(i));
        for (size_t i = 0; i < num_bot; ++i)
            emit(indices(matrix.bottom_monomials(i)), matrix.bottom_size(i), matrix.bottom_coeffs(i));

        gamba_cuda_step_in step{};
        step.n_top          = static_cast<uint32_t>(num_top);
        step.n_bot          = static_cast<uint32_t>(num_bot);
        step.n_cols         = static_cast<uint32_t>(num_cols);
        step.n_coeff_arrays = static_cast<uint32_t>(cf_ptr.size());
        step.nnz            = col.size();
        step.row_ptr        = row_ptr.data();
        step.col_idx        = col.data();
        step.coeff_id       = coeff_id.data();
        step.coeff_ptr      = reinterpret_cast<void const* const*>(cf_ptr.data());
        step.coeff_len      = cf_len.data();
        step.flags          = GAMBA_CUDA_STEP_FINAL;

        gamba_cuda_step_out out{};
        if (gamba_cuda_reduce(m_handle, &step, &out) != 0)
            throw std::runtime_error(std::string("gamba_cuda_reduce: ") + gamba_cuda_last_error());
        if (out.n_new != num_bot)
            throw std::runtime_error("final inter-reduction returned a different number of rows");

        /* one output row per bottom row, same order (source/basis.hpp:664-683): bottom rows are sorted by descending lead
         * column (source/matrix.hpp:733-739), the engine returns ascending pivots, so result row r is bottom row
         * num_bot-1-r. Memory comes from basis_type::allocate_polynomial because import_matrix_row reinterprets the
         * index array in place as monomials and the basis frees it (source/basis.hpp:386-408, 780-801, 243-247);
         * absolute (not relative) column indices: m_col_to_mon was not partitioned (source/basis.hpp:681). */
        auto const* const cfs = static_cast<coeff_type const*>(out.coeff);
        matrix.m_new_rows.clear();
        matrix.m_new_coefs.clear();
        for (size_t i = 0; i < num_bot; ++i)
        {
            size_t const r   = num_bot - 1 - i;
            size_t const nnz = static_cast<size_t>(out.row_ptr[r + 1] - out.row_ptr[r]);
            auto const [new_mon, new_cfs] = basis_type::allocate_polynomial(nnz);
            auto* const new_ind           = reinterpret_cast<index_type*>(new_mon);
            for (size_t k = 0; k < nnz; ++k)
            {
                uint64_t const e = out.row_ptr[r] + k;
                new_ind[k]       = nonpiv_col[out.col_rel[e]];
                new_cfs[k]       = cfs[e];
            }
            matrix.m_new_rows.emplace_back(new_ind, nnz);
            matrix.m_new_coefs.emplace_back(new_cfs, nnz);
        }

        stats.linalg_reduce_walltime += out.seconds_total;
        stats.linalg_reduce_cputime += out.seconds_total;
    }

private:
    gamba_cuda_engine* m_handle{nullptr};
};

}  // namespace gamba
