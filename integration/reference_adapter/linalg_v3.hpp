// Reference-side binding #1: `source/linalg_v3.hpp` for gblanco92/GamBa on top of gamba-b200's C ABI.
//
// The reference's own linalg_v3.hpp is withheld (README.md:85, .gitignore:17 of the reference); its open driver only
// needs a class with this interface. Drop this file into the reference's source/ (or put this directory on the include
// path in front of it), add -I<gamba-b200>/include, link libgamba_cuda.so. It is what oracle/build_ref.sh compiles
// against the reference's untouched sources to prove the boundary (tests/test_ref_adapter_*.py).
//
//   linalg_v3<C,O> engine{basis.field, params.seed}       source/f4.hpp:66     -> gamba_cuda_create (Montgomery form)
//   engine.initialize(matrix)                              source/f4.hpp:91     -> flatten matrix_f4 into gamba_cuda_step_in
//   engine.reduce(params) -> duration<double>              source/f4.hpp:93     -> gamba_cuda_reduce, repack for extract_new_rows
//   members read by matrix_f4::extract_new_rows            source/matrix.hpp:799-895
#pragma once
#include <chrono>
#include <cstdint>
#include <format>
#include <iostream>
#include <span>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "config.hpp"
#include "field.hpp"
#include "gamba_cuda.h"
#include "params.hpp"
#include "stats.hpp"

namespace gamba
{

template <class CoefficientType, class MonomialOrder>
class matrix_f4;   // matrix.hpp includes this header before the class is complete (source/matrix.hpp:28)

template <class CoefficientType, class MonomialOrder>
class linalg_v3
{
public:
    using coeff_type = CoefficientType;
    using index_type = uint32_t;
    /* wider than coeff_type: extract_new_rows asserts row_ptr[...] < max(coeff_type) (source/matrix.hpp:855-856) */
    using row_type = uint64_t;

    linalg_v3(field_traits<coeff_type> const& field, size_t const seed) : m_field{field}
    {
        gamba_cuda_config cfg{};
        cfg.p           = static_cast<uint32_t>(field.n);
        cfg.coeff_bytes = sizeof(coeff_type);
        cfg.repr        = GAMBA_CUDA_REPR_MONTGOMERY;   /* the driver keeps Montgomery residues, source/field.hpp:41-51 */
        cfg.seed        = seed;
        cfg.world_size  = 1;
        if (gamba_cuda_abi_version() != GAMBA_CUDA_ABI_VERSION)
            throw std::runtime_error("libgamba_cuda: ABI version mismatch");
        if (gamba_cuda_create(&cfg, &m_handle) != 0)
            throw std::runtime_error(std::string("gamba_cuda_create: ") + gamba_cuda_last_error());
        /* basis coefficient arrays are immutable until basis.clear() (source/reduce.hpp:54), which happens after the
         * last call on this engine: keep them on the device across rounds */
        gamba_cuda_set_option(m_handle, "coeff_cache", 1);
    }

    linalg_v3(linalg_v3 const&)            = delete;
    linalg_v3& operator=(linalg_v3 const&) = delete;

    ~linalg_v3() { gamba_cuda_destroy(m_handle); }

    /* source/f4.hpp:91 — monomial handles -> block-form columns (pivot columns first), shared coefficient arrays by id */
    void initialize(matrix_f4<CoefficientType, MonomialOrder>& matrix)
    {
        size_t const num_cols = matrix.num_cols();
        size_t const num_top  = matrix.m_top_rows.size();
        size_t const num_bot  = matrix.m_bottom_rows.size();

        m_is_pivot.assign(num_cols, 0);
        for (auto const& row : matrix.m_top_rows)
            m_is_pivot[row[0].data().idx] = 1;

        /* top rows come sorted by ascending lead column (source/matrix.hpp:634-639): pivot column number = row number */
        m_block_col.resize(num_cols);
        uint32_t next_piv = 0, next_nonpiv = static_cast<uint32_t>(num_top);
        for (size_t c = 0; c < num_cols; ++c)
            m_block_col[c] = m_is_pivot[c] ? next_piv++ : next_nonpiv++;

        m_num_nonpiv_cols = num_cols - num_top;
        m_num_bottom_rows = num_bot;

        m_row_ptr.assign(1, 0);
        m_col.clear();
        m_coeff_id.clear();
        m_cf_ptr.clear();
        m_cf_len.clear();
        m_ids.clear();

        auto emit = [this](auto const& row, coeff_type const* const cfs) {
            for (auto const mon : row)
                m_col.push_back(m_block_col[mon.data().idx]);
            m_row_ptr.push_back(m_col.size());
            /* all multiples of one generator share its coefficient array (source/matrix.hpp:189-193) */
            auto const [it, fresh] = m_ids.try_emplace(cfs, static_cast<uint32_t>(m_cf_ptr.size()));
            if (fresh)
            {
                m_cf_ptr.push_back(cfs);
                m_cf_len.push_back(static_cast<uint32_t>(row.size()));
            }
            m_coeff_id.push_back(it->second);
        };
        for (size_t i = 0; i < num_top; ++i)
            emit(matrix.m_top_rows[i], matrix.m_top_coeffs[i]);
        for (size_t i = 0; i < num_bot; ++i)
            emit(matrix.m_bottom_rows[i], matrix.m_bottom_coeffs[i]);

        m_step                = gamba_cuda_step_in{};
        m_step.n_top          = static_cast<uint32_t>(num_top);
        m_step.n_bot          = static_cast<uint32_t>(num_bot);
        m_step.n_cols         = static_cast<uint32_t>(num_cols);
        m_step.n_coeff_arrays = static_cast<uint32_t>(m_cf_ptr.size());
        m_step.nnz            = m_col.size();
        m_step.row_ptr        = m_row_ptr.data();
        m_step.col_idx        = m_col.data();
        m_step.coeff_id       = m_coeff_id.data();
        m_step.coeff_ptr      = reinterpret_cast<void const* const*>(m_cf_ptr.data());
        m_step.coeff_len      = m_cf_len.data();
    }

    /* source/f4.hpp:93 */
    std::chrono::duration<double> reduce(gamba_params const& params)
    {
        gamba_cuda_step_out out{};
        if (gamba_cuda_reduce(m_handle, &m_step, &out) != 0)
            throw std::runtime_error(std::string("gamba_cuda_reduce: ") + gamba_cuda_last_error());

        /* CSR (ascending pivot = decreasing leading monomial, the order source/basis.hpp:725-727 relies on) -> the
         * column-major blocks of NUM_ROWS_BUFFER rows that extract_new_rows reads: element (row r, relative non-pivot
         * column j) at [r + j * NUM_ROWS_BUFFER] (source/matrix.hpp:851-858) */
        size_t const num_new  = out.n_new;
        size_t const num_bufs = (num_new + NUM_ROWS_BUFFER - 1) / NUM_ROWS_BUFFER;
        auto const* const cfs = static_cast<coeff_type const*>(out.coeff);

        m_newrows_buffers.clear();
        m_newrows_piv_ind.clear();
        m_newrows_num_nnz.assign(num_bufs * NUM_ROWS_BUFFER, 0);
        m_newrows_is_pivot.assign(num_bufs * NUM_ROWS_BUFFER, 0);
        for (size_t b = 0; b < num_bufs; ++b)
        {
            size_t const rows_here = std::min<size_t>(NUM_ROWS_BUFFER, num_new - b * NUM_ROWS_BUFFER);
            m_newrows_buffers.emplace_back(static_cast<size_t>(NUM_ROWS_BUFFER) * m_num_nonpiv_cols, row_type{0});
            /* freed by the caller with delete[] (source/matrix.hpp:879) */
            m_newrows_piv_ind.emplace_back(new index_type[rows_here], rows_here);
        }
        for (size_t r = 0; r < num_new; ++r)
        {
            size_t const b = r / NUM_ROWS_BUFFER, k = r % NUM_ROWS_BUFFER;
            m_newrows_piv_ind[b][k] = out.piv_rel[r];
            m_newrows_is_pivot[r]   = 1;
            /* the pivot term is not counted (source/matrix.hpp:832-834) */
            m_newrows_num_nnz[r] = static_cast<size_t>(out.row_ptr[r + 1] - out.row_ptr[r] - 1);
            for (uint64_t e = out.row_ptr[r]; e < out.row_ptr[r + 1]; ++e)
                m_newrows_buffers[b][k + static_cast<size_t>(out.col_rel[e]) * NUM_ROWS_BUFFER] = cfs[e];
        }

        stats.zero_reductions += out.zero_reductions;
        stats.linalg_walltime += out.seconds_total;
        stats.linalg_cputime += out.seconds_total;
        if (params.rounds_info)
            std::cout << std::format("{:>9}", num_new) << std::flush;   /* the "new gens" cell of the round table */

        return std::chrono::duration<double>(out.seconds_total);
    }

    /* read by matrix_f4::extract_new_rows (source/matrix.hpp:802-885) */
    std::vector<uint8_t> m_is_pivot{};
    std::vector<std::vector<row_type>> m_newrows_buffers{};
    std::vector<std::span<index_type>> m_newrows_piv_ind{};
    std::vector<size_t> m_newrows_num_nnz{};
    std::vector<uint8_t> m_newrows_is_pivot{};
    size_t m_num_nonpiv_cols{0};
    size_t m_num_bottom_rows{0};

private:
    field_traits<coeff_type> const m_field;
    gamba_cuda_engine* m_handle{nullptr};
    gamba_cuda_step_in m_step{};
    std::vector<uint32_t> m_block_col{};
    std::vector<uint64_t> m_row_ptr{};
    std::vector<uint32_t> m_col{}, m_coeff_id{}, m_cf_len{};
    std::vector<coeff_type const*> m_cf_ptr{};
    std::unordered_map<coeff_type const*, uint32_t> m_ids{};
};

}  // namespace gamba
