// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_engine.hpp).
//
// Exact, deterministic, single-threaded. Stage 1 eliminates every pivot column
// of every row to reduce, left to right, with a dense 64-bit accumulator and
// delayed reduction (what the reference's kernel names saxpy / reduce_u32 /
// modular_u32 suggest, CMakeLists.txt:81-90). Stage 2 brings D' to canonical
// reduced row echelon form (exact; the reference's Monte-Carlo variant yields the
// same row space with high probability, README.md:5).
#include "oracle_engine.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>

namespace gb {

static uint32_t inv_mod_u32(uint32_t a, uint32_t p) {
    int64_t t = 0, nt = 1, r = p, nr = a % p;
    while (nr) { int64_t q = r / nr, x = t - q * nt; t = nt; nt = x; x = r - q * nr; r = nr; nr = x; }
    return (uint32_t)(t < 0 ? t + p : t);
}

void OracleEngine::reduce_step(const StepMatrix& m, NewRows& out) {
    if (threads_ > 1) { oracle_reduce_step_mt(m, out, threads_); return; }
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t p = m.p, ntop = m.n_top, nbot = m.n_bot, ncols = m.n_cols, nnp = ncols - ntop;
    const uint64_t P2 = (uint64_t)p * p;
    out = NewRows{};

    // ---- stage 1: C|D -> 0|D' ------------------------------------------------
    std::vector<uint64_t> acc(ncols);
    std::vector<std::vector<uint32_t>> rows;  // non-zero reduced rows, dense over non-pivot columns
    uint64_t fma_top = 0;
    for (uint32_t r = 0; r < nbot; ++r) {
        std::fill(acc.begin(), acc.end(), 0ull);
        const uint32_t row = ntop + r;
        const uint32_t* cf = m.cf_ptr[m.coeff_id[row]];
        uint32_t lo = ncols;
        for (uint64_t j = 0, n = m.row_ptr[row + 1] - m.row_ptr[row]; j < n; ++j) {
            const uint32_t c = m.col_at(row, j);
            acc[c] = cf[j];
            if (c < lo) lo = c;
        }
        for (uint32_t c = lo; c < ntop; ++c) {
            uint64_t v = acc[c] % p;
            if (!v) continue;
            const uint64_t mult = p - v;  // row -= v * pivot_row(c)
            const uint32_t* pc = m.cf_ptr[m.coeff_id[c]];
            const uint64_t a = m.row_ptr[c], b = m.row_ptr[c + 1];
            fma_top += b - a;
            for (uint64_t j = 1; j < b - a; ++j) {
                const uint32_t cc = m.col_at(c, j);
                uint64_t x = acc[cc] + mult * pc[j];
                acc[cc] = x >= P2 ? x - P2 : x;
            }
        }
        std::vector<uint32_t> d(nnp);
        bool nz = false;
        for (uint32_t j = 0; j < nnp; ++j) { d[j] = (uint32_t)(acc[ntop + j] % p); nz |= d[j] != 0; }
        if (nz) rows.push_back(std::move(d));
    }

    // ---- stage 2: canonical RREF of D' ---------------------------------------
    std::vector<int32_t> piv_row(nnp, -1);
    std::vector<std::vector<uint32_t>> piv;  // monic pivot rows
    std::vector<uint32_t> piv_col;
    std::vector<uint64_t> w(nnp);
    auto axpy_from = [&](uint32_t from, uint64_t mult, const std::vector<uint32_t>& src) {
        for (uint32_t j = from; j < nnp; ++j) {
            uint64_t x = w[j] + mult * src[j];
            w[j] = x >= P2 ? x - P2 : x;
        }
    };
    for (auto& d : rows) {
        for (uint32_t j = 0; j < nnp; ++j) w[j] = d[j];
        int32_t lead = -1;
        for (uint32_t j = 0; j < nnp; ++j) {
            uint64_t v = w[j] % p; w[j] = v;
            if (!v) continue;
            if (piv_row[j] >= 0) { axpy_from(j, p - v, piv[piv_row[j]]); w[j] = 0; }
            else if (lead < 0) lead = (int32_t)j;
        }
        if (lead < 0) continue;
        uint32_t inv = inv_mod_u32((uint32_t)w[lead], p);
        std::vector<uint32_t> nr(nnp);
        for (uint32_t j = 0; j < nnp; ++j) nr[j] = (uint32_t)((w[j] % p) * inv % p);
        piv_row[lead] = (int32_t)piv.size();
        piv.push_back(std::move(nr)); piv_col.push_back((uint32_t)lead);
    }
    // back substitution, highest pivot first
    std::vector<uint32_t> ord(piv.size());
    for (uint32_t i = 0; i < ord.size(); ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return piv_col[a] > piv_col[b]; });
    for (uint32_t i : ord) {
        auto& row = piv[i];
        for (uint32_t j = 0; j < nnp; ++j) w[j] = row[j];
        for (uint32_t j = piv_col[i] + 1; j < nnp; ++j) {
            uint64_t v = w[j] % p; w[j] = v;
            if (v && piv_row[j] >= 0) { axpy_from(j, p - v, piv[piv_row[j]]); w[j] = 0; }
        }
        for (uint32_t j = 0; j < nnp; ++j) row[j] = (uint32_t)(w[j] % p);
    }
    // emit by ascending pivot column
    std::reverse(ord.begin(), ord.end());
    out.n_new = (uint32_t)ord.size();
    out.row_ptr.assign(1, 0);
    for (uint32_t i : ord) {
        for (uint32_t j = piv_col[i]; j < nnp; ++j)
            if (piv[i][j]) { out.col.push_back(j); out.val.push_back(piv[i][j]); }
        out.row_ptr.push_back(out.col.size());
    }
    out.fma_top = fma_top;
    out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace gb
