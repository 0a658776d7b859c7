// Adapter: the host driver's Engine interface on top of the C ABI
// (include/gamba_cuda.h). This is the product path: it has no CPU fallback —
// if the CUDA engine cannot be created the driver stops with the error.
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/gamba_cuda.h"
#include "f4.hpp"

namespace gb {

class CudaEngine final : public Engine {
public:
    CudaEngine(const Params& prm, uint32_t p) {
        gamba_cuda_config cfg{};
        cfg.p = p; cfg.coeff_bytes = 4; cfg.repr = GAMBA_CUDA_REPR_STANDARD;
        cfg.device = 0; cfg.seed = prm.seed; cfg.rank = 0; cfg.world_size = 1;
        if (gamba_cuda_create(&cfg, &e_)) throw std::runtime_error(std::string("gamba_cuda_create: ") + gamba_cuda_last_error());
        // the basis' coefficient arrays are immutable for the life of the run: keep them on the device
        if (gamba_cuda_set_option(e_, "coeff_cache", 1)) throw std::runtime_error(gamba_cuda_last_error());
        if (!std::getenv("GAMBA_NO_PINNED")) {                // step matrices are built in page-locked memory
            host_alloc_hook().alloc = gamba_cuda_host_alloc;
            host_alloc_hook().release = gamba_cuda_host_free;
        }
        if (const char* o = std::getenv("GAMBA_OPTIONS")) {   // "key=value,key=value" -> gamba_cuda_set_option
            std::string s(o);
            for (size_t a = 0; a < s.size();) {
                size_t b = s.find(',', a); if (b == std::string::npos) b = s.size();
                const std::string kv = s.substr(a, b - a);
                const size_t eq = kv.find('=');
                if (eq != std::string::npos && gamba_cuda_set_option(e_, kv.substr(0, eq).c_str(), std::stoll(kv.substr(eq + 1))))
                    throw std::runtime_error(std::string("GAMBA_OPTIONS: ") + gamba_cuda_last_error());
                a = b + 1;
            }
        }
    }
    ~CudaEngine() override {
        if (std::getenv("GAMBA_DEBUG_TIMERS"))
            std::fprintf(stderr, "dbg engine: call %.3f (stage %.3f) copy-out %.3f | device: eliminate %.3f echelon %.3f\n", t_call_, t_stage_, t_copy_, t_elim_, t_ech_);
        gamba_cuda_destroy(e_);
    }
    const char* name() const override { return "cuda-b200"; }
    void reduce_step(const StepMatrix& m, NewRows& out) override { run(m, out, 0); }
    void reduce_final(const StepMatrix& m, NewRows& out) override { run(m, out, GAMBA_CUDA_STEP_FINAL); }

private:
    gamba_cuda_engine* e_ = nullptr;
    double t_call_ = 0, t_stage_ = 0, t_copy_ = 0, t_elim_ = 0, t_ech_ = 0;
    std::vector<const void*> ptrs_;
    void run(const StepMatrix& m, NewRows& out, uint32_t flags) {
        gamba_cuda_step_in in{};
        in.n_top = m.n_top; in.n_bot = m.n_bot; in.n_cols = m.n_cols;
        in.n_coeff_arrays = (uint32_t)m.cf_ptr.size(); in.nnz = m.nnz();
        in.row_ptr = m.row_ptr.data(); in.col_idx = m.col.data(); in.coeff_id = m.coeff_id.data();
        if (m.src) {     // indirect form: no O(nnz) rewrite on the host
            in.col_idx = m.src; in.col_idx_len = m.src_len; in.row_src = m.row_src.data();
            in.col_map = m.col_map.data(); in.col_map_len = (uint32_t)m.col_map.size();
        }
        ptrs_.assign(m.cf_ptr.begin(), m.cf_ptr.end());
        in.coeff_ptr = ptrs_.data(); in.coeff_len = m.cf_len.data(); in.flags = flags;
        gamba_cuda_step_out o{};
        const double t0 = now_s();
        if (gamba_cuda_reduce(e_, &in, &o)) throw std::runtime_error(std::string("gamba_cuda_reduce: ") + gamba_cuda_last_error());
        const double t1 = now_s();
        gamba_cuda_counters c{};
        gamba_cuda_get_counters(e_, &c);
        t_call_ += t1 - t0; t_stage_ += c.seconds_stage_last;
        out = NewRows{};
        out.n_new = o.n_new;
        out.row_ptr.assign(o.row_ptr, o.row_ptr + o.n_new + 1);
        out.col.assign(o.col_rel, o.col_rel + o.nnz_new);
        const uint32_t* v = static_cast<const uint32_t*>(o.coeff);
        out.val.assign(v, v + o.nnz_new);
        out.seconds = o.seconds_total; out.kernel_seconds = o.seconds_device; out.fma_top = c.fma_top;
        t_elim_ += o.seconds_eliminate; t_ech_ += o.seconds_echelon;
        char note[160];
        static const char* const mode[] = {"sparse", "sparse-blocked", "tc-solve"};
        std::snprintf(note, sizeof note, "[%s T=%u | stage %.1f elim %.1f (operands %.1f) echelon %.1f ms | mc %u]", mode[c.solve_mode % 3], c.tile_cols,
                      (o.seconds_device - o.seconds_eliminate - o.seconds_echelon) * 1e3, o.seconds_eliminate * 1e3, c.seconds_operands_last * 1e3,
                      o.seconds_echelon * 1e3, c.mc_ech_combinations);
        out.note = note;
        t_copy_ += now_s() - t1;
    }
};

Engine* gb_make_engine(const Params& prm, uint32_t p) { return new CudaEngine(prm, p); }

}  // namespace gb
