// Adapter: the host driver's Engine interface on top of the C ABI
// (include/gamba_cuda.h). This is the product path: it has no CPU fallback —
// if the CUDA engine cannot be created the driver stops with the error.
#include <dlfcn.h>

#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>

#include "../../include/gamba_cuda.h"
#include "f4.hpp"

namespace gb {

// ---- N GPUs of one box, one process: NCCL loaded at run time (only when --gpus > 1) -----------------------------
// The C ABI never names NCCL (include/gamba_cuda.h: two callbacks on device pointers and a stream); this driver wires
// them to ncclBroadcast / ncclAllGather the way INTEGRATION.md tells a C++ host to. Prototypes as in nccl.h (2.x).
struct Nccl {
    typedef void* comm_t;
    int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* lib = nullptr;
    static constexpr int kUint8 = 1;        // ncclUint8
    void load() {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) throw std::runtime_error(std::string("--gpus > 1 needs NCCL: ") + dlerror());
        auto sym = [&](const char* n) { void* f = dlsym(lib, n); if (!f) throw std::runtime_error(std::string("NCCL symbol missing: ") + n); return f; };
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(sym("ncclBroadcast"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    }
};

struct RankCtx { Nccl* nccl; Nccl::comm_t comm; };

static int nccl_bcast_cb(void* user, void* buf, size_t bytes, int root, void* stream) {
    auto* c = static_cast<RankCtx*>(user);
    return c->nccl->Broadcast(buf, buf, bytes, Nccl::kUint8, root, c->comm, stream);
}
static int nccl_gather_cb(void* user, const void* send, void* recv, size_t bytes, void* stream) {
    auto* c = static_cast<RankCtx*>(user);
    return c->nccl->AllGather(send, recv, bytes, Nccl::kUint8, c->comm, stream);
}


class CudaEngine final : public Engine, public RowGenDevice {
public:
    // ---- RowGenDevice: thin wrappers of gamba_cuda_rowgen_* on the engine of GPU 0 ----
    RowGenDevice* rowgen() override { return std::getenv("GAMBA_HOST_ROWGEN") ? nullptr : this; }
    void begin(uint32_t stride) override { if (gamba_cuda_rowgen_begin(e_, stride)) fail("gamba_cuda_rowgen_begin"); }
    uint64_t add_monomials(uint32_t count, const uint16_t* ex, const uint32_t* hv) override {
        uint64_t first = 0;
        if (gamba_cuda_rowgen_add_monomials(e_, count, ex, hv, &first)) fail("gamba_cuda_rowgen_add_monomials");
        return first;
    }
    void add_polys(uint32_t n_polys, const uint32_t* n_terms, const uint32_t* const* mon, uint64_t* first) override {
        if (gamba_cuda_rowgen_add_polys(e_, n_polys, n_terms, mon, first)) fail("gamba_cuda_rowgen_add_polys");
    }
    uint32_t add_rows(uint32_t n_rows, const uint64_t* first_term, const uint32_t* n_terms, const uint16_t* q_ex,
                      const uint32_t* q_hash, uint32_t* lead_out) override {
        uint32_t n = 0;
        if (gamba_cuda_rowgen_add_rows(e_, n_rows, first_term, n_terms, q_ex, q_hash, lead_out, &n)) fail("gamba_cuda_rowgen_add_rows");
        return n;
    }
    void fetch(uint32_t first, uint32_t count, uint16_t* ex_out, uint32_t* hv_out) override {
        if (gamba_cuda_rowgen_fetch_monomials(e_, first, count, ex_out, hv_out)) fail("gamba_cuda_rowgen_fetch_monomials");
    }
    const uint32_t* rows_device(uint64_t& n_entries) override {
        const uint32_t* p = nullptr;
        if (gamba_cuda_rowgen_rows(e_, &p, &n_entries)) fail("gamba_cuda_rowgen_rows");
        return p;
    }
    void download_rows(uint64_t first, uint64_t count, uint32_t* out) override {
        if (gamba_cuda_rowgen_download_rows(e_, first, count, out)) fail("gamba_cuda_rowgen_download_rows");
    }
    [[noreturn]] static void fail(const char* what) { throw std::runtime_error(std::string(what) + ": " + gamba_cuda_last_error()); }

    CudaEngine(const Params& prm, uint32_t p) : world_(std::max(1, prm.gpus)) {
        gamba_cuda_config cfg{};
        cfg.p = p; cfg.coeff_bytes = 4; cfg.repr = GAMBA_CUDA_REPR_STANDARD;
        cfg.device = 0; cfg.seed = prm.seed; cfg.rank = 0; cfg.world_size = (uint32_t)world_;
        if (gamba_cuda_create(&cfg, &e_)) throw std::runtime_error(std::string("gamba_cuda_create: ") + gamba_cuda_last_error());
        // the basis' coefficient arrays are immutable for the life of the run: keep them on the device
        if (gamba_cuda_set_option(e_, "coeff_cache", 1)) throw std::runtime_error(gamba_cuda_last_error());
        if (world_ > 1) {
            // rows of C|D sharded r % world over the GPUs of the box: this thread drives rank 0 and hands the
            // step over; one thread per other GPU calls gamba_cuda_reduce(e, NULL, ...) and receives it by ncclBroadcast
            nccl_.load();
            std::vector<int> devs(world_);
            for (int r = 0; r < world_; ++r) devs[r] = r;
            comms_.resize(world_);
            if (int rc = nccl_.CommInitAll(comms_.data(), world_, devs.data())) throw std::runtime_error(std::string("ncclCommInitAll: ") + nccl_.GetErrorString(rc));
            ctx_.resize(world_); peers_.assign(world_, nullptr); peers_[0] = e_;
            for (int r = 0; r < world_; ++r) ctx_[r] = RankCtx{&nccl_, comms_[r]};
            for (int r = 1; r < world_; ++r) {
                cfg.device = (uint32_t)r; cfg.rank = (uint32_t)r;
                if (gamba_cuda_create(&cfg, &peers_[r])) throw std::runtime_error(std::string("gamba_cuda_create (GPU ") + std::to_string(r) + "): " + gamba_cuda_last_error());
            }
            for (int r = 0; r < world_; ++r) {
                gamba_cuda_set_broadcast(peers_[r], nccl_bcast_cb, &ctx_[r]);
                gamba_cuda_set_allgather(peers_[r], nccl_gather_cb, &ctx_[r]);
            }
            for (int r = 1; r < world_; ++r) workers_.emplace_back([this, r] { worker(r); });
        }
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
                for (int r = 0; r < world_ && eq != std::string::npos; ++r)
                    if (gamba_cuda_set_option(r ? peers_[r] : e_, kv.substr(0, eq).c_str(), std::stoll(kv.substr(eq + 1))))
                        throw std::runtime_error(std::string("GAMBA_OPTIONS: ") + gamba_cuda_last_error());
                a = b + 1;
            }
        }
    }
    ~CudaEngine() override {
        if (std::getenv("GAMBA_DEBUG_TIMERS"))
            std::fprintf(stderr, "dbg engine: call %.3f (stage %.3f) copy-out %.3f | device: eliminate %.3f echelon %.3f\n", t_call_, t_stage_, t_copy_, t_elim_, t_ech_);
        if (world_ > 1) {
            { std::lock_guard<std::mutex> g(mu_); quit_ = true; ++gen_; }
            cv_.notify_all();
            for (auto& t : workers_) t.join();
            for (int r = 1; r < world_; ++r) gamba_cuda_destroy(peers_[r]);
        }
        gamba_cuda_destroy(e_);
        for (auto c : comms_) if (c) nccl_.CommDestroy(c);
    }
    const char* name() const override { return world_ > 1 ? "cuda-b200-multi" : "cuda-b200"; }
    void reduce_step(const StepMatrix& m, NewRows& out) override { run(m, out, 0); }
    void reduce_final(const StepMatrix& m, NewRows& out) override { run(m, out, GAMBA_CUDA_STEP_FINAL); }

private:
    // ranks 1 .. world-1: wait for a step, take part in it (the step arrives over NVLink), drop the replicated result
    void worker(int r) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (quit_) return;
            }
            gamba_cuda_step_out o{};
            const int rc = gamba_cuda_reduce(peers_[r], nullptr, &o);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (rc) { failed_ = true; err_ = gamba_cuda_last_error(); }
                ++done_;
            }
            cv_done_.notify_one();
        }
    }
    int world_ = 1;
    Nccl nccl_;
    std::vector<Nccl::comm_t> comms_;
    std::vector<RankCtx> ctx_;
    std::vector<gamba_cuda_engine*> peers_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, cv_done_;
    uint64_t gen_ = 0;
    int done_ = 0;
    bool quit_ = false, failed_ = false;
    std::string err_;
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
        in.coeff_ptr = ptrs_.data(); in.coeff_len = m.cf_len.data(); in.flags = flags | (m.src_on_device ? GAMBA_CUDA_STEP_DEVICE_COLIDX : 0u);
        gamba_cuda_step_out o{};
        const double t0 = now_s();
        if (world_ > 1) { { std::lock_guard<std::mutex> g(mu_); done_ = 0; ++gen_; } cv_.notify_all(); }
        const int rc0 = gamba_cuda_reduce(e_, &in, &o);
        const std::string err0 = rc0 ? gamba_cuda_last_error() : "";
        if (world_ > 1) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_done_.wait(lk, [&] { return done_ == world_ - 1; });
            if (failed_) throw std::runtime_error("gamba_cuda_reduce on a peer GPU: " + err_);
        }
        if (rc0) throw std::runtime_error("gamba_cuda_reduce: " + err0);
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
