// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_engine.hpp).
// Plain-C entry point so tests/ and bench.py can run the oracle on a step matrix
// through ctypes. Same argument meaning as gamba_cuda_step_in (include/gamba_cuda.h).
#include <cstdlib>
#include <cstring>

#include "oracle_engine.hpp"

extern "C" {

struct gamba_oracle_result {
    uint32_t n_new;
    uint64_t nnz_new;
    uint64_t* row_ptr;   // n_new + 1
    uint32_t* col_rel;   // nnz_new
    uint32_t* coeff;     // nnz_new
    uint64_t fma_top;
    double seconds;
};

// threads <= 1: single-thread timing mode; threads > 1: checker mode (oracle_checker.cpp)
int gamba_oracle_reduce_mt(uint32_t p, uint32_t n_top, uint32_t n_bot, uint32_t n_cols,
                           const uint64_t* row_ptr, const uint32_t* col_idx, const uint32_t* coeff_id,
                           uint32_t n_coeff_arrays, const uint32_t* const* coeff_ptr, const uint32_t* coeff_len,
                           uint32_t threads, gamba_oracle_result* res) {
    try {
        gb::StepMatrix m;
        m.p = p; m.n_top = n_top; m.n_bot = n_bot; m.n_cols = n_cols;
        const uint32_t nr = n_top + n_bot;
        m.row_ptr.assign(row_ptr, row_ptr + nr + 1);
        m.col.assign(col_idx, col_idx + row_ptr[nr]);
        m.coeff_id.assign(coeff_id, coeff_id + nr);
        m.cf_ptr.assign(coeff_ptr, coeff_ptr + n_coeff_arrays);
        m.cf_len.assign(coeff_len, coeff_len + n_coeff_arrays);
        gb::OracleEngine eng(threads); gb::NewRows out;
        eng.reduce_step(m, out);
        res->n_new = out.n_new; res->nnz_new = out.col.size();
        res->row_ptr = (uint64_t*)std::malloc(sizeof(uint64_t) * (out.n_new + 1));
        res->col_rel = (uint32_t*)std::malloc(sizeof(uint32_t) * (out.col.size() + 1));
        res->coeff = (uint32_t*)std::malloc(sizeof(uint32_t) * (out.col.size() + 1));
        std::memcpy(res->row_ptr, out.row_ptr.data(), sizeof(uint64_t) * (out.n_new + 1));
        std::memcpy(res->col_rel, out.col.data(), sizeof(uint32_t) * out.col.size());
        std::memcpy(res->coeff, out.val.data(), sizeof(uint32_t) * out.val.size());
        res->fma_top = out.fma_top; res->seconds = out.seconds;
        return 0;
    } catch (...) { return 1; }
}

int gamba_oracle_reduce(uint32_t p, uint32_t n_top, uint32_t n_bot, uint32_t n_cols,
                        const uint64_t* row_ptr, const uint32_t* col_idx, const uint32_t* coeff_id,
                        uint32_t n_coeff_arrays, const uint32_t* const* coeff_ptr, const uint32_t* coeff_len,
                        gamba_oracle_result* res) {
    return gamba_oracle_reduce_mt(p, n_top, n_bot, n_cols, row_ptr, col_idx, coeff_id, n_coeff_arrays, coeff_ptr, coeff_len, 1, res);
}

void gamba_oracle_free(gamba_oracle_result* res) {
    std::free(res->row_ptr); std::free(res->col_rel); std::free(res->coeff);
    res->row_ptr = nullptr; res->col_rel = nullptr; res->coeff = nullptr;
}

}  // extern "C"
