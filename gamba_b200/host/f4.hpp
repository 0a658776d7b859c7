// Host-side F4 driver: S-pair selection, symbolic preprocessing, matrix
// assembly and basis bookkeeping around the linear-algebra boundary.
//
// Keeps the role of the reference's open driver (source/f4.hpp:47-130,
// source/spair.hpp:190-264, source/update.hpp:29-306, source/matrix.hpp:215-673,
// source/basis.hpp:287-380,697-769, source/reduce.hpp:27-77) so that
// `gamba -i in -o out` is a drop-in; the code is an independent implementation
// (one basis table + one per-step table, flat row arenas, standard residues).
#pragma once
#include <chrono>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "engine.hpp"
#include "io.hpp"
#include "mono.hpp"
#include "par.hpp"

namespace gb {

struct Params {
    bool all_pairs = false;
    long max_spairs = 2000;     // 0 = all pairs of minimal degree
    int num_elim = 0;
    bool no_reduce = false;
    uint64_t seed = 967557673ull;
    int verbose = 2;
    int threads = 0;            // -t: host threads (0 = all hardware threads, at most 16)
    int gpus = 1;               // --gpus: GPUs of this box the linear algebra is sharded over (C|D rows r % gpus)
    std::string dump_dir;       // write every step matrix (+ result) here
    long dump_min_nnz = 0;
    long dump_top = 0;          // keep only the N steps with the most non-zeros (0 = all)
    long dump_every = 0;        // dump only every N-th step (0/1 = all)
    long stop_after = 0;        // debugging: leave the F4 loop after this many steps (0 = run to the end)
};

struct Stats {
    double t_update = 0, t_select = 0, t_rows = 0, t_symbolic = 0, t_convert = 0,
           t_linalg = 0, t_linalg_kernel = 0, t_insert = 0, t_final = 0, t_final_linalg = 0, t_total = 0;
    uint64_t steps = 0, pairs_reduced = 0, gm_removed = 0, rows_reduced = 0, zero_reductions = 0;
    uint64_t nnz_in = 0, fma_top = 0, max_rows = 0, max_cols = 0, max_nnz = 0;
};

class F4 {
public:
    F4(const PolySystem& in, const Params& prm, Engine& eng);
    void run();
    void export_basis(PolySystem& out) const;
    Stats stats;

private:
    struct Poly {
        std::vector<uint32_t> mon;  // basis-table handles, decreasing
        std::vector<uint32_t> cf;   // residues in [0,p), cf[0] == 1
        // term data laid out contiguously for row generation (filled on first use as a reducer or pair row):
        // streaming reads instead of two random accesses into the basis table per term
        std::vector<uint16_t> ex;   // [terms][stride]
        std::vector<uint32_t> hv;   // [terms]
        uint32_t deg = 0;
        uint64_t lm_mask = 0;
        bool redundant = false;
        uint64_t dev_first = ~0ull;  // where its terms start in the device's term store (row generation on the GPU)
    };
    struct Pair { uint32_t i, j, lcm; int32_t deg; };

    const PolySystem& in_;
    Params prm_;
    Engine& eng_;
    uint32_t p_;
    MonoCtx cx_;
    MonoTable bt_, st_;
    std::vector<Poly> G_;
    std::vector<Pair> P_;
    bool trivial_ = false;
    int nthreads_ = 1;          // host threads for row generation, reducer search, column mapping
    std::unique_ptr<ThreadPool> pool_;
    double dbg_t_[4] = {0, 0, 0, 0};
    long dbg_waves_ = 0;
    double upd_t_[5] = {0, 0, 0, 0, 0};
    std::vector<uint16_t> upd_lcm_;      // update(): lcms of the candidate pairs of one generator

    // per-step state
    std::vector<uint32_t> row_mon_;      // flat step-table handles
    std::vector<uint64_t> row_off_;
    std::vector<uint32_t> row_gen_;
    std::vector<uint8_t> row_top_;
    std::vector<uint32_t> row_lead_;     // step handle of every row's leading monomial
    RowGenDevice* rg_ = nullptr;         // rows are generated on the engine's device (CUDA engine) instead of here
    uint32_t bt_dev_n_ = 0;              // basis-table monomials already uploaded to it
    std::vector<uint8_t> has_red_;       // per step-table monomial
    std::vector<uint32_t> order_;        // column -> step handle
    std::vector<uint32_t> nonpiv_col_;   // relative non-pivot index -> column
    StepMatrix M_;
    NewRows R_;
    std::vector<std::pair<uint64_t, std::string>> dump_kept_;   // --dump-top: (nnz, tag) of the steps on disk

    void import_generators();
    void update(uint32_t h);
    void update_batch(size_t from);
    void update_all(size_t from);
    long select_pairs(int32_t& deg);
    void clear_step();
    void add_rows(const std::vector<uint32_t>& gens, const std::vector<uint16_t>& qs, const std::vector<uint32_t>& hqs,
                  const std::vector<uint8_t>& tops);
    void build_pair_rows(long n_sel);
    void symbolic();
    void convert();
    void insert_new_rows();
    void final_reduce();
    void make_trivial();
    void dump_step(const char* tag);
    std::vector<uint32_t> nonredundant() const;
};

inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace gb
