// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's two closed linear-algebra engines
//   linalg_v3<Coeff,Order>::initialize/reduce   (call sites source/f4.hpp:91-93;
//                                                output read by source/matrix.hpp:799-895)
//   linalg<Matrix>::reduce                      (call site source/reduce.hpp:51-52)
// against the boundary contract of SURVEY.md §8(b). The reference's own source
// for this path (source/linalg_v3.hpp, source/linalg.hpp, source/kernel/*) is
// withheld (reference README.md:85, .gitignore:16-19), so this is a restatement
// of the *contract*, not of code.
//
// Pinning: end to end only — the host driver + this engine reproduces the
// reference's MAGMA-generated golden bases (test/results/*.out.txt.tar.gz; the
// committed subset is tests/golden/) byte for byte, in both --max-spairs modes.
// Per-step D' block: PARITY UNPINNED against the reference (its per-step form is
// unobservable); the CUDA path is compared bit-exact against this oracle's
// canonical reduced row echelon form instead.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may use anything in this directory. The product never does.
#pragma once
#include "../gamba_b200/host/engine.hpp"

namespace gb {

// Checker mode (oracle_checker.cpp): same contract, all host threads; a different stage-2 algorithm on purpose.
void oracle_reduce_step_mt(const StepMatrix& m, NewRows& out, unsigned n_threads);

class OracleEngine final : public Engine {
public:
    // threads <= 1: the single-thread timing mode (what cpu_baseline measures; the reference is single-threaded,
    // README.md:9,71); threads > 1: the checker mode
    explicit OracleEngine(unsigned threads = 1) : threads_(threads) {}
    const char* name() const override { return threads_ > 1 ? "oracle-cpu-mt" : "oracle-cpu"; }
    void reduce_step(const StepMatrix& m, NewRows& out) override;

private:
    unsigned threads_;
};

}  // namespace gb
