// `gamba -i in.txt -o out.txt [...]` — same CLI, input grammar and output layout
// as the reference binary (source/main.cpp:32-73 options, :83-163 flow).
// The linear-algebra engine is whatever gb_make_engine() links in: the CUDA
// engine behind the C ABI for the product binary, the CPU oracle for
// oracle/_build/gamba_oracle (checker / CPU baseline only).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>

#include "f4.hpp"

namespace gb { Engine* gb_make_engine(const Params& prm, uint32_t p); }

static void usage() {
    std::puts(
        "GamBa-B200: a Groebner basis engine (F4, B200 linear algebra)\n"
        "Usage: gamba [OPTIONS]\n"
        "  -i,--input-file TEXT REQUIRED   Input filename\n"
        "  -o,--output-file TEXT           Output filename\n"
        "  -e,--num-elim UINT              Number of variables in the first elimination block\n"
        "  --max-spairs UINT               Max pairs of minimal degree per round (0 = all)\n"
        "  --all-pairs                     Select all pairs in the queue in each round\n"
        "  --no-reduce                     Do not compute a reduced Groebner basis\n"
        "  -t,--threads UINT               Host threads for the symbolic phases (0 = all, at most 16)\n"
        "  --gpus UINT                     GPUs of this box to shard the linear algebra over (default 1)\n"
        "  -v,--verbose UINT               Verbosity level (default 2)\n"
        "  --dump-steps DIR                Write every step matrix and its result to DIR\n"
        "  --dump-min-nnz N                ... only steps with at least N non-zeros\n"
        "  --dump-top N                    ... only the N steps with the most non-zeros\n"
        "  --dump-every N                  ... only every N-th step\n"
        "  --stop-after N                  (debugging) leave the F4 loop after N steps");
}

int main(int argc, char** argv) {
    std::string input, output;
    gb::Params prm;
    auto need = [&](int& i, const char* opt) -> std::string {
        std::string a = argv[i];
        size_t eq = a.find('=');
        if (eq != std::string::npos && a.compare(0, 2, "--") == 0) return a.substr(eq + 1);
        if (i + 1 >= argc) { std::fprintf(stderr, "%s: 1 required argument missing\n", opt); std::exit(106); }
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        std::string key = a.substr(0, a.find('='));
        try {
            if (key == "-i" || key == "--input-file") input = need(i, "--input-file");
            else if (key == "-o" || key == "--output-file") output = need(i, "--output-file");
            else if (key == "-e" || key == "--num-elim") prm.num_elim = std::stoi(need(i, "--num-elim"));
            else if (key == "--max-spairs") prm.max_spairs = std::stol(need(i, "--max-spairs"));
            else if (key == "--all-pairs") prm.all_pairs = true;
            else if (key == "--no-reduce") prm.no_reduce = true;
            else if (key == "-t" || key == "--threads") prm.threads = std::stoi(need(i, "--threads"));
            else if (key == "--gpus") prm.gpus = std::stoi(need(i, "--gpus"));
            else if (key == "-v" || key == "--verbose") prm.verbose = std::stoi(need(i, "--verbose"));
            else if (key == "--dump-steps") prm.dump_dir = need(i, "--dump-steps");
            else if (key == "--dump-min-nnz") prm.dump_min_nnz = std::stol(need(i, "--dump-min-nnz"));
            else if (key == "--dump-top") prm.dump_top = std::stol(need(i, "--dump-top"));
            else if (key == "--dump-every") prm.dump_every = std::stol(need(i, "--dump-every"));
            else if (key == "--stop-after") prm.stop_after = std::stol(need(i, "--stop-after"));
            else if (key == "-h" || key == "--help") { usage(); return 0; }
            else { std::fprintf(stderr, "The following argument was not expected: %s\n", a.c_str()); return 109; }
        } catch (std::exception const&) {
            std::fprintf(stderr, "Could not convert argument of %s\n", key.c_str());
            return 104;
        }
    }
    if (input.empty()) { std::fprintf(stderr, "--input-file is required\n"); return 106; }
    if (prm.max_spairs < 0) { std::fprintf(stderr, "--max-spairs: value out of range\n"); return 105; }

    const auto t_start = std::chrono::steady_clock::now();
    auto since_start = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    double t_setup = 0, t_run_end = 0;
    try {
        std::ifstream in(input);
        if (in.fail()) throw std::runtime_error("Error opening input file.");
        gb::PolySystem sys;
        sys.read(in);
        gb::PolySystem outsys;
        outsys.vars = sys.vars; outsys.p = sys.p;
        if (sys.p != 0) {
            // Neither the engine nor the F4 state is destroyed on the way out: the process ends below with _Exit, and the OS and
            // the CUDA driver give tens of GB of rows, hash tables, page-locked buffers and mapped device memory back in one go
            // (running the destructors took ~10 s of a 23 s kat15-15 run). Error paths unwind normally.
            gb::Engine* eng = gb::gb_make_engine(prm, sys.p);
            if (prm.verbose > 0)
                std::printf("GamBa-B200 [engine = %s, seed = %llu]\n", eng->name(), (unsigned long long)prm.seed);
            gb::F4* f4p = nullptr;
            t_setup = since_start();
            try { f4p = new gb::F4(sys, prm, *eng); f4p->run(); t_run_end = since_start(); f4p->export_basis(outsys); }
            catch (...) { delete f4p; delete eng; throw; }
            gb::F4& f4 = *f4p;
            if (prm.verbose > 0) {
                auto const& s = f4.stats;
                std::printf("------------------------------------------------------------\n");
                std::printf("update %.3f  select %.3f  rows %.3f  symbolic %.3f  convert %.3f  insert %.3f\n",
                            s.t_update, s.t_select, s.t_rows, s.t_symbolic, s.t_convert, s.t_insert);
                std::printf("linalg (wall)       %12.3f sec   [device-only %.3f]\n", s.t_linalg, s.t_linalg_kernel);
                std::printf("reduce basis        %12.3f sec   [linalg %.3f]\n", s.t_final, s.t_final_linalg);
                std::printf("overall (wall)      %12.3f sec\n", s.t_total);
                std::printf("steps %llu  pairs %llu  gm-removed %llu  rows-reduced %llu  zero-reductions %llu\n",
                            (unsigned long long)s.steps, (unsigned long long)s.pairs_reduced,
                            (unsigned long long)s.gm_removed, (unsigned long long)s.rows_reduced,
                            (unsigned long long)s.zero_reductions);
                std::printf("nnz_in %llu  fma_top %llu  max matrix %llu x %llu (%llu nnz)\n",
                            (unsigned long long)s.nnz_in, (unsigned long long)s.fma_top,
                            (unsigned long long)s.max_rows, (unsigned long long)s.max_cols,
                            (unsigned long long)s.max_nnz);
                if (s.t_linalg > 0)
                    std::printf("linalg nnz reduced/s %.4g\n", (double)s.nnz_in / s.t_linalg);
            }
        }
        if (!output.empty()) {
            std::ofstream out(output);
            if (out.fail()) throw std::runtime_error("Error opening output file.");
            outsys.write(out);
            out.close();
            if (out.fail()) throw std::runtime_error("Error writing output file.");
        }
        if (prm.verbose > 0 && sys.p != 0)
            std::printf("process: setup (input, engine) %.3f  run %.3f  export + output file %.3f sec\n", t_setup, t_run_end - t_setup, since_start() - t_run_end);
        std::fflush(stdout);
        std::fflush(stderr);
        std::_Exit(0);
    } catch (std::exception const& e) {
        std::cerr << e.what() << std::endl;
        return -1;
    } catch (...) {
        std::cerr << "Unexpected error." << std::endl;
        return -2;
    }
    return 0;
}
