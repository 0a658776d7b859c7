// ORACLE — TEST INFRASTRUCTURE ONLY. The C ABI of include/gamba_cuda.h implemented on the CPU oracle, so that the
// reference-side adapters (integration/reference_adapter/) can be checked end to end in a container without a GPU:
// oracle/_ref/gamba_ref_cpu = the reference's open driver + the adapters + this library. It exists to test the ADAPTERS
// (block-form flattening, Montgomery form in and out, NUM_ROWS_BUFFER repack, allocate_polynomial ownership); the
// product never loads it, and it is not a fallback of libgamba_cuda.so (different file name, only linked by build_ref.sh).
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gamba_cuda.h"
#include "../oracle_engine.hpp"

static thread_local std::string g_err;

struct gamba_cuda_engine {
    gamba_cuda_config cfg{};
    uint32_t r_mod_p = 1, rinv_mod_p = 1;
    gamba_cuda_counters ctr{};
    std::vector<uint64_t> row_ptr;
    std::vector<uint32_t> col, piv, v32;
    std::vector<uint16_t> v16;
    std::vector<uint8_t> v8;
};

static uint32_t inv_mod(uint32_t a, uint32_t p) {
    int64_t t = 0, nt = 1, r = p, nr = a % p;
    while (nr) { int64_t q = r / nr, x = t - q * nt; t = nt; nt = x; x = r - q * nr; r = nr; nr = x; }
    return (uint32_t)(t < 0 ? t + p : t);
}

extern "C" {

int gamba_cuda_abi_version(void) { return GAMBA_CUDA_ABI_VERSION; }
const char* gamba_cuda_last_error(void) { return g_err.c_str(); }

int gamba_cuda_create(const gamba_cuda_config* cfg, gamba_cuda_engine** out) {
    if (!cfg || !out || cfg->p < 3 || (cfg->coeff_bytes != 1 && cfg->coeff_bytes != 2 && cfg->coeff_bytes != 4)) { g_err = "bad configuration"; return 1; }
    auto* e = new gamba_cuda_engine();
    e->cfg = *cfg;
    if (cfg->repr == GAMBA_CUDA_REPR_MONTGOMERY) {   // R = 2^32 for 1-/2-byte coefficients, 2^64 for 4-byte ones (field.hpp:41-51)
        uint64_t r = 1;
        for (uint32_t i = 0; i < (cfg->coeff_bytes == 4 ? 64u : 32u); ++i) r = r * 2 % cfg->p;
        e->r_mod_p = (uint32_t)r; e->rinv_mod_p = inv_mod((uint32_t)r, cfg->p);
    }
    *out = e;
    return 0;
}
void gamba_cuda_destroy(gamba_cuda_engine* e) { delete e; }
void* gamba_cuda_host_alloc(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
void gamba_cuda_host_free(void* p) { std::free(p); }
int gamba_cuda_set_option(gamba_cuda_engine*, const char*, int64_t) { return 0; }
int gamba_cuda_get_counters(gamba_cuda_engine* e, gamba_cuda_counters* c) { *c = e->ctr; return 0; }
int gamba_cuda_set_allgather(gamba_cuda_engine*, gamba_cuda_allgather_fn, void*) { return 0; }
int gamba_cuda_set_broadcast(gamba_cuda_engine*, gamba_cuda_broadcast_fn, void*) { return 0; }
int gamba_cuda_stage(gamba_cuda_engine*, const gamba_cuda_step_in*) { g_err = "not implemented by the oracle shim"; return 1; }
int gamba_cuda_reduce_staged(gamba_cuda_engine*, gamba_cuda_step_out*) { g_err = "not implemented by the oracle shim"; return 1; }

int gamba_cuda_reduce(gamba_cuda_engine* e, const gamba_cuda_step_in* in, gamba_cuda_step_out* out) {
    try {
        const uint32_t p = e->cfg.p, nr = in->n_top + in->n_bot;
        gb::StepMatrix m;
        m.p = p; m.n_top = in->n_top; m.n_bot = in->n_bot; m.n_cols = in->n_cols;
        m.row_ptr.assign(in->row_ptr, in->row_ptr + nr + 1);
        m.col.resize(in->nnz);
        for (uint32_t r = 0; r < nr; ++r) {
            const uint64_t len = in->row_ptr[r + 1] - in->row_ptr[r], src = in->row_src ? in->row_src[r] : in->row_ptr[r];
            for (uint64_t j = 0; j < len; ++j) {
                const uint32_t h = in->col_idx[src + j];
                m.col[in->row_ptr[r] + j] = in->col_map ? in->col_map[h] : h;
            }
        }
        m.coeff_id.assign(in->coeff_id, in->coeff_id + nr);
        std::vector<std::vector<uint32_t>> pool(in->n_coeff_arrays);
        for (uint32_t i = 0; i < in->n_coeff_arrays; ++i) {
            pool[i].resize(in->coeff_len[i]);
            for (uint32_t k = 0; k < in->coeff_len[i]; ++k) {
                uint64_t v = e->cfg.coeff_bytes == 4 ? ((const uint32_t*)in->coeff_ptr[i])[k]
                           : e->cfg.coeff_bytes == 2 ? ((const uint16_t*)in->coeff_ptr[i])[k] : ((const uint8_t*)in->coeff_ptr[i])[k];
                pool[i][k] = (uint32_t)(v % p * e->rinv_mod_p % p);          // Montgomery -> standard (identity for REPR_STANDARD)
            }
            m.cf_ptr.push_back(pool[i].data()); m.cf_len.push_back(in->coeff_len[i]);
        }
        const char* t = std::getenv("GAMBA_ORACLE_THREADS");
        gb::OracleEngine eng(t ? (unsigned)std::atoi(t) : 1u);
        gb::NewRows res;
        eng.reduce_step(m, res);
        if ((in->flags & GAMBA_CUDA_STEP_FINAL) && res.n_new != in->n_bot) throw std::runtime_error("final inter-reduction: rank differs from the number of rows");
        e->row_ptr = res.row_ptr; e->col = res.col;
        e->piv.resize(res.n_new);
        for (uint32_t r = 0; r < res.n_new; ++r) e->piv[r] = res.col[res.row_ptr[r]];
        const size_t nnz = res.val.size();
        e->v32.resize(nnz);
        for (size_t i = 0; i < nnz; ++i) e->v32[i] = (uint32_t)((uint64_t)res.val[i] * e->r_mod_p % p);   // standard -> Montgomery
        *out = gamba_cuda_step_out{};
        out->n_new = res.n_new; out->n_nonpiv = in->n_cols - in->n_top; out->nnz_new = nnz;
        out->row_ptr = e->row_ptr.data(); out->col_rel = e->col.data(); out->piv_rel = e->piv.data();
        if (e->cfg.coeff_bytes == 4) out->coeff = e->v32.data();
        else if (e->cfg.coeff_bytes == 2) { e->v16.assign(e->v32.begin(), e->v32.end()); out->coeff = e->v16.data(); }
        else { e->v8.assign(e->v32.begin(), e->v32.end()); out->coeff = e->v8.data(); }
        out->zero_reductions = in->n_bot - res.n_new;
        out->seconds_total = res.seconds;
        e->ctr.steps += 1; e->ctr.fma_top = res.fma_top;
        return 0;
    } catch (std::exception const& ex) { g_err = ex.what(); return 1; }
}

}  // extern "C"
