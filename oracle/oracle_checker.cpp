// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_engine.hpp).
//
// Checker mode of the CPU oracle: the same contract as OracleEngine::reduce_step (canonical reduced row
// echelon form of D' = D - C A^-1 B), computed with all the host threads the box has so that
// the big named systems (cyclic9, kat15, eco15: hours on one thread) can be checked in seconds per step.
//   stage 1  rows to reduce are independent given A|B: a dynamic parallel-for over the rows,
//            each thread with its own dense 64-bit accumulator - the loop body is the one of the timing mode;
//   stage 2  right-looking Gauss-Jordan on the non-zero rows of D': per pivot (smallest lead column first) the
//            pivot row is made monic and eliminated from every other row that has an entry in its column,
//            rows dealt to the threads, two barriers per pivot. A different algorithm from the timing mode's
//            row-by-row reduction on purpose: the reduced row echelon form of a row space is unique, so the two
//            must agree bit for bit (tests/test_oracle_cpu.py checks that they do).
// The single-thread timing mode (oracle_engine.cpp) stays what cpu_baseline measures.
#include <algorithm>
#include <atomic>
#include <barrier>
#include <chrono>
#include <thread>

#include "oracle_engine.hpp"

namespace gb {

static uint32_t chk_inv(uint32_t a, uint32_t p) {
    int64_t t = 0, nt = 1, r = p, nr = a % p;
    while (nr) { int64_t q = r / nr, x = t - q * nt; t = nt; nt = x; x = r - q * nr; r = nr; nr = x; }
    return (uint32_t)(t < 0 ? t + p : t);
}

// x mod p for x < 2^63 (Barrett with a 64-bit reciprocal; the plain % is a 64-bit division per entry)
struct Barrett {
    uint64_t p, mu;
    explicit Barrett(uint32_t p_) : p(p_), mu(~0ull / p_) {}
    uint32_t operator()(uint64_t x) const {
        const uint64_t q = (uint64_t)(((unsigned __int128)x * mu) >> 64);
        uint64_t r = x - q * p;
        if (r >= p) r -= p;
        if (r >= p) r -= p;
        return (uint32_t)r;
    }
};

void oracle_reduce_step_mt(const StepMatrix& m, NewRows& out, unsigned n_threads) {
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t p = m.p, ntop = m.n_top, nbot = m.n_bot, ncols = m.n_cols, nnp = ncols - ntop;
    const uint64_t P2 = (uint64_t)p * p;
    n_threads = std::max(1u, std::min(n_threads, 256u));
    const Barrett red(p);
    out = NewRows{};

    // ---- stage 1 ----------------------------------------------------------------
    std::vector<std::vector<uint32_t>> drow(nbot);      // reduced rows over the non-pivot columns (empty = zero row)
    std::atomic<uint32_t> next{0};
    std::atomic<uint64_t> fma_top{0};
    auto stage1 = [&]() {
        std::vector<uint64_t> acc(ncols);
        uint64_t fma = 0;
        for (;;) {
            const uint32_t r = next.fetch_add(1);
            if (r >= nbot) break;
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
                const uint64_t v = acc[c] % p;
                if (!v) continue;
                const uint64_t mult = p - v;
                const uint32_t* pc = m.cf_ptr[m.coeff_id[c]];
                const uint64_t a = m.row_ptr[c], b = m.row_ptr[c + 1];
                fma += b - a;
                for (uint64_t j = 1; j < b - a; ++j) {
                    const uint32_t cc = m.col_at(c, j);
                    const uint64_t x = acc[cc] + mult * pc[j];
                    acc[cc] = x >= P2 ? x - P2 : x;
                }
            }
            bool nz = false;
            for (uint32_t j = 0; j < nnp && !nz; ++j) nz = acc[ntop + j] % p != 0;
            if (nz) {
                std::vector<uint32_t> d(nnp);
                for (uint32_t j = 0; j < nnp; ++j) d[j] = (uint32_t)(acc[ntop + j] % p);
                drow[r] = std::move(d);
            }
        }
        fma_top += fma;
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < n_threads; ++t) th.emplace_back(stage1);
        stage1();
        for (auto& x : th) x.join();
    }

    // ---- stage 2: Gauss-Jordan, rows dealt to the threads -------------------------
    std::vector<std::vector<uint32_t>> rows;
    for (auto& d : drow) if (!d.empty()) rows.push_back(std::move(d));
    const uint32_t nr = (uint32_t)rows.size();
    constexpr uint32_t NONE = 0xffffffffu;
    std::vector<uint32_t> lead(nr, NONE);     // lead column of a row that is not a pivot row yet (NONE: zero or pivot row)
    std::vector<uint8_t> is_piv(nr, 0);
    std::vector<uint32_t> piv_rows, piv_cols;
    for (uint32_t i = 0; i < nr; ++i)
        for (uint32_t j = 0; j < nnp; ++j) if (rows[i][j]) { lead[i] = j; break; }
    struct Best { uint32_t col, row; };
    std::vector<Best> local(n_threads, Best{NONE, NONE});
    uint32_t cur_row = NONE, cur_col = NONE;
    bool done = false;
    std::barrier sync((std::ptrdiff_t)n_threads);
    auto stage2 = [&](unsigned tid) {
        for (;;) {
            Best b{NONE, NONE};
            for (uint32_t i = tid; i < nr; i += n_threads)
                if (lead[i] < b.col) b = Best{lead[i], i};       // ties: smallest row index (strict <, ascending i)
            local[tid] = b;
            sync.arrive_and_wait();
            if (tid == 0) {
                Best g{NONE, NONE};
                for (auto const& x : local) if (x.col < g.col || (x.col == g.col && x.row < g.row)) g = x;
                if (g.col == NONE) done = true;
                else {
                    auto& pr = rows[g.row];
                    const uint64_t inv = chk_inv(pr[g.col], p);
                    for (uint32_t j = g.col; j < nnp; ++j) pr[j] = red(pr[j] * inv);
                    is_piv[g.row] = 1; lead[g.row] = NONE;
                    piv_rows.push_back(g.row); piv_cols.push_back(g.col);
                    cur_row = g.row; cur_col = g.col;
                }
            }
            sync.arrive_and_wait();
            if (done) break;
            const uint32_t c = cur_col;
            const uint32_t* pr = rows[cur_row].data();
            for (uint32_t i = tid; i < nr; i += n_threads) {
                if (i == cur_row) continue;
                uint32_t* w = rows[i].data();
                const uint64_t v = w[c];
                if (!v) continue;
                const uint64_t mult = p - v;
                for (uint32_t j = c; j < nnp; ++j) w[j] = red(w[j] + mult * pr[j]);
                if (!is_piv[i]) {                               // its lead was at c or further right
                    uint32_t l = lead[i];
                    while (l < nnp && !w[l]) ++l;
                    lead[i] = l < nnp ? l : NONE;
                }
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < n_threads; ++t) th.emplace_back(stage2, t);
        stage2(0);
        for (auto& x : th) x.join();
    }
    std::vector<uint32_t> ord(piv_rows.size());
    for (uint32_t i = 0; i < ord.size(); ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return piv_cols[a] < piv_cols[b]; });
    out.n_new = (uint32_t)ord.size();
    out.row_ptr.assign(1, 0);
    for (uint32_t i : ord) {
        auto const& row = rows[piv_rows[i]];
        for (uint32_t j = piv_cols[i]; j < nnp; ++j)
            if (row[j]) { out.col.push_back(j); out.val.push_back(row[j]); }
        out.row_ptr.push_back(out.col.size());
    }
    out.fma_top = fma_top.load();
    out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace gb
