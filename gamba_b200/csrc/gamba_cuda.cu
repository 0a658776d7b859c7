// Host side of the B200 engine and the C ABI of include/gamba_cuda.h.
//
// Role of the reference's closed linalg_v3<Coeff,Order> / linalg<Matrix> classes
// (call sites source/f4.hpp:63-100, source/reduce.hpp:51-52): stage one step matrix
// in HBM, eliminate C|D by A|B, echelonise D', hand the new rows back.
// There is no CPU fallback in this file: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_scan.cuh>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gamba_cuda.h"
#include "field.cuh"
#include "kernels_echelon.cuh"
#include "kernels_echelon_mma.cuh"
#include "kernels_elim.cuh"
#include "kernels_elim_block.cuh"
#include "kernels_mc.cuh"
#include "kernels_prep.cuh"
#include "kernels_rowgen.cuh"
#include "kernels_schur.cuh"
#include "kernels_schur_tc.cuh"
#include "kernels_tri.cuh"
#include "layout.cuh"

namespace gcu {

#define GCU_CHECK(expr)                                                                         \
    do {                                                                                        \
        cudaError_t err__ = (expr);                                                             \
        if (err__ != cudaSuccess)                                                               \
            throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(err__));    \
    } while (0)

static thread_local std::string g_last_error;

// Device buffers come from a stream-ordered memory pool PRIVATE to the engine (release threshold "never": a freed block stays in
// the pool for the next step): growing a buffer is a pool operation ordered on the engine's stream, not a cudaFree (which
// synchronises the device) plus a cudaMalloc (milliseconds per GB) - 2 s of a kat15-15 run before. Matrices grow from step to step,
// so freed blocks of the wrong size pile up in the pool: when an allocation fails the stream is drained, the pool trimmed and the
// allocation retried. The pool is destroyed with the engine, so nothing stays reserved on the device afterwards and no other
// user of the device's default pool is affected.
static thread_local cudaStream_t g_alloc_stream = nullptr;
static thread_local cudaMemPool_t g_alloc_pool = nullptr;

static void* pool_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t err = cudaMallocFromPoolAsync(&p, bytes, g_alloc_pool, g_alloc_stream);
    if (err == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        GCU_CHECK(cudaStreamSynchronize(g_alloc_stream));
        GCU_CHECK(cudaMemPoolTrimTo(g_alloc_pool, 0));
        err = cudaMallocFromPoolAsync(&p, bytes, g_alloc_pool, g_alloc_stream);
    }
    if (err != cudaSuccess) {
        cudaGetLastError();
        throw std::runtime_error("out of device memory allocating " + std::to_string(bytes >> 20) + " MiB: " + cudaGetErrorString(err));
    }
    return p;
}

// Big buffers (>= 32 MB) do not come from the pool: matrices grow from step to step, a pool block cannot be extended, and the
// freed smaller blocks pile up (reserved memory ~2-3x the live size, then an out-of-memory trim, then everything again from the
// driver at ~25 ms per GB: stalls of 0.2-1.4 s inside kat15-15 steps). They are virtual-memory ranges instead: a 256 GB
// address reservation per buffer, physical memory mapped at its end as the buffer grows - contents stay, nothing is copied or
// freed, and the driver is asked for every byte exactly once. Driver entry points through the runtime (no libcuda link).
struct VmmApi {
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};
static VmmApi& vmm() {
    static VmmApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && *fn && q == cudaDriverEntryPointSuccess;
        };
        api.ok = get("cuMemAddressReserve", (void**)&api.AddressReserve) && get("cuMemAddressFree", (void**)&api.AddressFree) &&
                 get("cuMemCreate", (void**)&api.Create) && get("cuMemRelease", (void**)&api.Release) && get("cuMemMap", (void**)&api.Map) &&
                 get("cuMemUnmap", (void**)&api.Unmap) && get("cuMemSetAccess", (void**)&api.SetAccess) &&
                 get("cuMemGetAllocationGranularity", (void**)&api.GetGranularity);
        if (!api.ok) cudaGetLastError();
        if (std::getenv("GAMBA_NO_VMM")) api.ok = false;
    }
    return api;
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    // virtual-memory mode
    CUdeviceptr va = 0;
    size_t va_size = 0;
    std::vector<CUmemGenericAllocationHandle> chunks;
    static constexpr size_t VMM_MIN = (size_t)32 << 20, VMM_RANGE = (size_t)256 << 30;

    // contents are kept when the buffer is in virtual-memory mode (va != 0), lost otherwise
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if ((va || bytes >= VMM_MIN) && vmm().ok && bytes <= VMM_RANGE) { grow_vmm(bytes); return; }
        release();
        const size_t want = bytes + std::min<size_t>(bytes / 2, (size_t)12 << 30) + 256;     // few, generous reallocations
        p = pool_alloc(want);
        cap = want;
    }
    void grow_vmm(size_t bytes) {
        VmmApi& v = vmm();
        int dev = 0;
        GCU_CHECK(cudaGetDevice(&dev));
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = dev;
        size_t gran = 0;
        if (v.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || !gran) gran = (size_t)2 << 20;
        if (!va) {
            if (p) { GCU_CHECK(cudaFreeAsync(p, g_alloc_stream)); p = nullptr; cap = 0; }      // was a small pool block: its contents go
            if (v.AddressReserve(&va, VMM_RANGE, 0, 0, 0) != CUDA_SUCCESS) throw std::runtime_error("cuMemAddressReserve failed");
            va_size = VMM_RANGE;
        }
        size_t want = bytes + std::min<size_t>(bytes / 4, (size_t)4 << 30);      // growing is cheap: little slack
        want = std::min((want + gran - 1) / gran * gran, va_size);
        const size_t delta = want - cap;
        CUmemGenericAllocationHandle h{};
        CUresult r = v.Create(&h, delta, &prop, 0);
        if (r == CUDA_ERROR_OUT_OF_MEMORY) {        // give the pool's idle blocks back to the driver and try again
            GCU_CHECK(cudaStreamSynchronize(g_alloc_stream));
            GCU_CHECK(cudaMemPoolTrimTo(g_alloc_pool, 0));
            r = v.Create(&h, delta, &prop, 0);
        }
        if (r != CUDA_SUCCESS) throw std::runtime_error("out of device memory growing a buffer to " + std::to_string(want >> 20) + " MiB");
        if (v.Map(va + cap, delta, 0, h, 0) != CUDA_SUCCESS) { v.Release(h); throw std::runtime_error("cuMemMap failed"); }
        CUmemAccessDesc acc{};
        acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (v.SetAccess(va + cap, delta, &acc, 1) != CUDA_SUCCESS) throw std::runtime_error("cuMemSetAccess failed");
        chunks.push_back(h);
        cap = want; p = reinterpret_cast<void*>(va);
    }
    void release() {
        if (va) {        // the owner has drained its stream (gamba_cuda_destroy)
            VmmApi& v = vmm();
            if (cap) v.Unmap(va, cap);
            for (auto h : chunks) v.Release(h);
            v.AddressFree(va, va_size);
            chunks.clear(); va = 0; va_size = 0;
        } else if (p) cudaFreeAsync(p, g_alloc_stream);
        p = nullptr; cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
};

struct HostBuf {  // pinned
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) GCU_CHECK(cudaFreeHost(p));
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 2 + 256;   // matrices grow from step to step: few, generous reallocations
        GCU_CHECK(cudaMallocHost(&p, want));
        cap = want;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
    ~HostBuf() { if (p) cudaFreeHost(p); }
};

static double wall_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// 2-D byte tensor [rows][cols] (cols contiguous) with a box of box_rows x box_cols and the 128-byte swizzle the
// tcgen05 operand descriptors of kernels_schur_tc.cuh expect. cuTensorMapEncodeTiled is a driver entry point:
// fetched through the runtime so that the library does not link libcuda.
static CUtensorMap make_tmap_u8(void* base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        GCU_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled is not available");
        encode = reinterpret_cast<encode_fn>(fn);
    }
    CUtensorMap m;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

}  // namespace gcu

struct gamba_cuda_engine {
    gamba_cuda_config cfg{};
    gcu::Fp f{};
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;
    cudaEvent_t ev[10]{};
    int max_smem_optin = 0, smem_per_sm = 0;
    int64_t opt_tile_cols = 0;
    int64_t opt_ctas_per_sm = 0;
    int64_t opt_elim_warps = 0;      // 4 or 8 warps per row to reduce (0 = automatic)
    int elim_warps = 8;
    int64_t opt_coeff_cache = 0;     // keep coefficient arrays on the device across steps (caller: immutable arrays)
    struct CacheEntry { uint64_t off; uint32_t len, first, last; };          // offset in d_pool, length, first and last coefficient
    std::unordered_map<const void*, CacheEntry> cache;                       // keyed by host pointer; an array whose length or end
                                                                             // coefficients changed (address reused) is uploaded again
    uint64_t cache_used = 0;
    int64_t opt_schur = 1;           // 1: D' = D + M B on the tensor cores when the limb planes fit, 0: sparse path
    int64_t opt_panel = 1;           // 1 (when it is predicted to pay) / 2 (always): blocked triangular solve on the tensor cores
                                     //    (kernels_tri.cuh): per pivot tile two tcgen05 products, no walk over the rows
    bool use_panel = false;
    uint32_t PANEL_T = 2048;         // pivot tile of the blocked solve: 128 * 2^j (the inverse diagonal tiles come from recursive doubling)
    int64_t opt_panel_t = 0;         // 0 = chosen per step from n_top
    int64_t opt_panel_max_top = 1000000;  // operands of K^2 / 2 bytes per limb: a cap besides the memory budget and the cost estimate
    gcu::DevBuf d_at, d_at_off, d_cd, d_cx, d_wr, d_vt, d_adt, d_gx, d_dinv;
    gcu::DevBuf d_bs_bitmap, d_bs_rank, d_bs_base, d_bs_klist, d_bs_op, d_bs_at;      // block-sparse operands (kernels_schur.cuh)
    std::vector<uint32_t> bs_tile_base;   // first block id of every pivot tile
    int64_t opt_bsparse = 1;         // 1: the products skip the blocks of A / B that hold no entry
    double last_occ = 1.0;           // occupied fraction of the blocks, previous step that measured it
    double t_inv_dev = 0;            // recursive doubling of the inverse diagonal tiles (part of the elimination time)
    gcu::DevBuf d_mc_s, d_mc_dt, d_mc_dc, d_mc_nz;
    int64_t opt_mc_ech = 1;          // Monte-Carlo combinations of the rows of D' in front of the echelon (accepted with a rank margin of 32)
    uint32_t mc_ech_k = 0;           // combinations used by the last echelon (0 = it ran on the rows themselves)
    uint64_t mc_ech_draw = 0;
    bool mc_fell_back = false;       // the last echelon could not allocate the combinations and ran on the rows
    std::vector<uint64_t> at_off;
    uint64_t at_bytes = 0;
    int64_t opt_block = 1;           // blocked sparse solve (kernels_elim_block.cuh): 0 never, 1 when the multipliers are dense, 2 always
    double last_density = 0;         // pivots hit per row / n_top of the previous step (predictor for this one)
    double last_a = 0, cur_a = 0;    // entries of A per pivot row (outside the diagonal blocks): previous step, staged step
    bool have_history = false;
    // Choice between the sparse kernel and the panel products: modelled times (rates measured with tools/step_stats.py).
    // The sparse kernel's work is  rows x hits x entries of the pivot rows hit: the pivot rows that are hit often are
    // longer than average (cyclic9: 4x, katsura: 1x) - hitw = applied entries / (rows x hits x mean length / 2), measured
    // whenever a step runs the sparse kernel (a ratio of counts: meaningful on small steps too), refreshed every 16 steps
    double hitw = 1;
    uint32_t since_explore = 0;
    // entries/s of the row-per-CTA kernel: 2.1e11 for p < 2^16, 1.3e11 above (two shared-memory REDs per entry)
    static double model_sparse(double nloc, double dens, double k, double a, bool small) { return nloc * dens * k * a * 0.5 / (small ? 2.1e11 : 1.3e11); }
    // blocked solve: products with the earlier tiles (K^2/2) and with the inverse diagonal tiles (K T) at the measured int8
    // rate, the inverse tiles themselves (K T^2 / 3, small batched products: a third of the rate), zeroing / scattering the
    // operands (bytes at ~3 TB/s), ~3 launches per tile
    // occ = occupied fraction of the 128 x bn blocks of A (block-sparse operands skip the rest)
    static double model_panel(double mpad, double k, double nl, double t, double occ) {
        const double rate = 1.2e15;
        return mpad * (k * k * 0.5 * occ + k * t) * nl * nl / rate + k * t * t / 3.0 * nl * nl / (rate / 3) +
               (nl * k * k * 0.5 * occ + 3.0 * nl * k * t) / 3e12 + (3.0 * k / t + 12) * 4e-6;
    }
    // sparse-blocked kernel: every block of R rows streams all of A: rows x nnz(A) row-entries at ~7e11 / s (kat15-15, K = 500 k)
    // (and never faster than one CTA streaming A once: ~6e8 entries/s, whatever the number of rows)
    static double model_block(double nloc, double k, double a, bool small) { return std::max(nloc * k * a / (small ? 7e11 : 4e11), k * a / 6e8); }
    double calib[3] = {1, 1, 1};     // measured / modelled time of the last steps in each mode (sparse, sparse-blocked, tc-solve)
    int64_t opt_panel_max_gb = 64;   // budget for the operands of the tc-solve alone (estimated with the block occupancy of the previous step)
    bool use_schur = false, use_block = false;   // decided per staged step (choose_layout)
    int64_t opt_schur_max_gb = 96;   // budget for the limb-plane operands of the products
    int64_t opt_schur_min_top = 1024; // fewer pivot rows: not worth the dense operands
    int64_t opt_schur_kchunk = 0;    // k-steps between accumulator folds (0 = the exactness bound; tests lower it)
    // Monte-Carlo combinations of the rows to reduce (kernels_mc.cuh): 0 never; 1 on steps with at least mc_min_top pivot rows when the
    // rank predicted from the previous step needs at most half as many combinations as there are rows; 2 on every step a blocked
    // solve can run on (64 combinations to start with when no rank is known)
    int64_t opt_mc = 1;
    int64_t opt_mc_k = 0;            // number of combinations (0 = automatic)
    int64_t opt_mc_min_top = 262144; // mc = 1: steps with fewer pivot rows eliminate the rows themselves (launch-bound: nothing to gain)
    int64_t opt_mc_min_rows = 256;   // mc = 1: ... and so do steps with fewer rows to reduce (with >= 262144 pivot rows the solve has a floor of
                                     // ~0.7 ms per pivot tile whatever the rows: 1013 rows of rank 13 against 453 k pivots took 728 ms in elim_kernel)
    int64_t last_rank = -1;          // rank of D' of the previous step (predictor for the number of combinations)
    uint32_t mc_plan = 0;            // combinations the staged step starts with (choose_layout; 0 = the rows themselves)
    uint32_t mc_k = 0;               // combinations the current elimination works on (0 = exact)
    uint32_t mc_base = 0;            // rows of D' that already hold reduced combinations of earlier draws of this step
    uint32_t mc_cap_rows = 0;        // rows D' must have room for (a later draw must not move the earlier ones)
    uint64_t mc_draw = 0;            // counter making every draw of multipliers different
    gcu::DevBuf d_mc_sel, d_mc_c;    // multipliers S [combinations][n_bot], dense C half of the combinations [combinations][n_tiles_piv * T]

    // staged step
    bool staged = false;
    gcu::Layout L{};
    uint32_t flags = 0;
    uint64_t nnz = 0;
    gcu::DevBuf d_row_ptr, d_col, d_coeff_id, d_coeff_off, d_pool, d_row_src, d_col_map;
    const uint64_t* s_row_ptr = nullptr;    // row pointers of the staged step on the device (ours, or the caller's with DEVICE_INDICES)
    gcu::DevBuf d_seg, d_ent, d_diag, d_dnz, d_first, d_scan_tmp;
    gcu::HostBuf h_pool, h_off;
    uint64_t h2d_bytes = 0;
    double t_stage_wall = 0, t_prep_dev = 0;

    // elimination / echelon state
    gcu::DevBuf d_dprime, d_row_nz, d_list_k, d_list_m, d_queue_stats;
    gcu::DevBuf d_em_live, d_em_xa, d_em_pbt;
    gcu::DevBuf d_mx, d_bt, d_mt;
    double t_sparse_dev = 0, t_schur_dev = 0;
    gcu::DevBuf d_lead, d_nz_list, d_piv_row, d_piv_col, d_is_piv, d_counts, d_out_ptr, d_out_col, d_out_val;
    uint32_t n_local = 0, n_local_max = 0, n_gather_rows = 0;
    bool eliminated = false;
    int elim_grid = 0;

    // results (host, engine-owned)
    std::vector<uint64_t> r_row_ptr;
    std::vector<uint32_t> r_col, r_piv, r_val32;
    std::vector<uint16_t> r_val16;
    std::vector<uint8_t> r_val8;
    gcu::HostBuf h_out_ptr, h_out_col, h_out_val, h_misc;

    gamba_cuda_allgather_fn gather_fn = nullptr;
    void* gather_user = nullptr;
    gamba_cuda_broadcast_fn bcast_fn = nullptr;
    void* bcast_user = nullptr;
    gcu::DevBuf d_hdr;
    gcu::HostBuf h_hdr;

    gamba_cuda_counters ctr{};
    double t_elim_dev = 0, t_ech_dev = 0;

    // row generation of symbolic preprocessing (kernels_rowgen.cuh)
    struct RowGenState {
        uint32_t stride = 0, sd = 0;
        gcu::DevBuf bex, bhv;              // the driver's basis monomials (exponents, hashes), appended as its table grows
        uint64_t n_basis = 0;
        gcu::DevBuf pmon;                  // term store of the basis polynomials: basis-monomial handles (append-only across steps)
        uint64_t terms = 0;
        gcu::DevBuf slots, hv, ex;         // the step's monomial table
        uint32_t n_slots = 0, cap = 0, count = 0;
        gcu::DevBuf rows;                  // handles of the step's rows, in the order the rows were added
        uint64_t n_rows_entries = 0;
        gcu::DevBuf cnt, b_first, b_off, b_q, b_hq, b_lead;
        gcu::HostBuf stage;
    } rg;
    void rg_begin(uint32_t stride);
    void rg_add_monomials(uint32_t count, const uint16_t* ex, const uint32_t* hv, uint64_t* first);
    void rg_add_polys(uint32_t n_polys, const uint32_t* n_terms, const uint32_t* const* mon, uint64_t* first);
    void rg_add_rows(uint32_t n_rows, const uint64_t* first_term, const uint32_t* n_terms, const uint16_t* q_ex, const uint32_t* q_hash,
                     uint32_t* lead_out, uint32_t* n_monomials);
    void rg_fetch(uint32_t first, uint32_t count, uint16_t* ex_out, uint32_t* hv_out);

    double dbg_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ~gamba_cuda_engine() {
        if (std::getenv("GAMBA_DEBUG_TIMERS"))
            std::fprintf(stderr, "dbg stage: hdr+layout %.3f alloc %.3f copies+pack %.3f alloc2 %.3f kernels+sync %.3f | reduce: elim-launch %.3f echelon+d2h %.3f\n",
                         dbg_[0], dbg_[1], dbg_[2], dbg_[3], dbg_[4], dbg_[5], dbg_[6]);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
    }
    // called by gamba_cuda_destroy after the buffers (members) are gone
    static void release_device(cudaStream_t st, cudaMemPool_t pl) {
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
        if (pl) cudaMemPoolDestroy(pl);
    }

    void choose_layout(uint32_t n_top, uint32_t n_bot, uint32_t n_cols, uint32_t step_flags, uint64_t step_nnz);
    void stage(const gamba_cuda_step_in* in);
    void eliminate();
    void exchange();
    void echelon(gamba_cuda_step_out* out, uint32_t mc_rows = 0);
    void run_reduce(gamba_cuda_step_out* out);
};

using namespace gcu;

void gamba_cuda_engine::choose_layout(uint32_t n_top, uint32_t n_bot, uint32_t n_cols, uint32_t step_flags, uint64_t step_nnz) {
    L = Layout{};
    L.n_top = n_top; L.n_bot = n_bot; L.n_cols = n_cols; L.n_nonpiv = n_cols - n_top; L.n_rows = n_top + n_bot;
    L.ld_dprime = (L.n_nonpiv + 3u) & ~3u;
    // Tile width from the number of rows (= warps) wanted in flight per SM: the kernel is
    // latency-bound, shared memory per SM / (warps per SM) is what one accumulator tile gets.
    const uint32_t gran = WINDOW;
    // 4 CTAs of 8 warps per SM. Measured with the Schur product on (tools/tune_ctas.sh, cyclic9-15 and kat14-15):
    // more resident rows with narrower tiles are slower (6: +10 %, 8: +25 %, 12-16 with 4 warps: 2x) - the
    // deferred passes then stream the pivot rows in pieces of a few entries per tile.
    uint32_t ctas = 4;
    elim_warps = ELIM_CTA_WARPS;
    if (opt_ctas_per_sm > 0) { ctas = (uint32_t)opt_ctas_per_sm; elim_warps = ctas > 8 ? 4 : ELIM_CTA_WARPS; }
    // fewer rows than SMs against many pivot rows (kat15-15: 13 rows, 410 k pivots, 88 % of them hit): a row is ONE CTA, and a warp
    // retires a 64-entry chunk per ~450 cycles of instruction latency whatever is in flight - so 32 warps share the row's stream
    if (opt_ctas_per_sm <= 0 && (uint64_t)(n_bot + std::max(1u, cfg.world_size) - 1) / std::max(1u, cfg.world_size) * 2 <= (uint64_t)prop.multiProcessorCount && n_top >= 65536)
        elim_warps = 32;
    if (opt_elim_warps == 2 || opt_elim_warps == 4 || opt_elim_warps == 8 || opt_elim_warps == 32) elim_warps = (int)opt_elim_warps;
    ctas = std::min<uint32_t>(ctas, 64u / elim_warps);
    const size_t per_cta = std::min<size_t>(((size_t)smem_per_sm / ctas) - 1024 - 64, (size_t)max_smem_optin - 64);
    const uint32_t words_per_col = f.small ? 1u : 2u;   // kernels_elim.cuh: Acc<SMALLP>::WORDS
    const size_t cta_words = per_cta / sizeof(uint32_t) - (size_t)elim_warps * ELIM_SCRATCH_WORDS;
    uint32_t tmax = (uint32_t)((cta_words - 4) / words_per_col - 1) / gran * gran;
    // Schur mode: the B half of the elimination is a dense product on the tensor cores (kernels_schur.cuh), and
    // (unless Monte-Carlo combinations are on) the A half runs R rows per CTA over narrower tiles
    {
        const uint32_t world = std::max(1u, cfg.world_size);
        const uint32_t nl = f.small ? 2u : 4u;
        const uint32_t sc_bm = 128, sc_bn = 128;     // the limb planes are padded to 128 rows (tile of either product kernel)
        const uint64_t kp = (n_top + SC_BK - 1) / SC_BK * SC_BK;
        const uint64_t nt_ = (L.n_nonpiv + sc_bn - 1) / sc_bn;
        // Three ways of finding the multipliers M = -C A^-1 (the A half of the elimination), chosen per step from modelled times:
        //   sparse          row-per-CTA kernel (kernels_elim.cuh): work = the pivot rows a row actually hits;
        //   sparse-blocked  R rows per CTA (kernels_elim_block.cuh): every entry of A is streamed once per block of rows - pays when
        //                   most pivot rows hit most rows (katsura / eco: 60-70 % of the multipliers are non-zero);
        //   tc-solve        blocked solve on the tensor cores (kernels_tri.cuh): n_bot K^2 / 2 dense multiply-adds whatever the
        //                   matrix holds; operands At_t [nl][T][t T] per tile (K^2/2 * nl bytes) + 3.25 * nl * K * T for the
        //                   inverse diagonal tiles - wins up to K ~ 170 k on katsura-like steps, by 1.5-3x on cyclic9.
        // The models use the previous step's multiplier density and entries per row of A (F4 steps change slowly); each is scaled
        // by a calibration factor = measured / modelled time of the steps that ran in that mode (run_reduce), so a model that is off
        // on a system corrects itself after one step in that mode; every 12th step the runner-up runs to keep its factor fresh.
        // With Monte-Carlo combinations (kernels_mc.cuh) the rows are dense: only the two blocked solves apply.
        if (opt_panel_t > 0) PANEL_T = (uint32_t)opt_panel_t;
        else PANEL_T = n_top >= 12288 ? 2048u : n_top >= 3072 ? 1024u : 512u;
        const double occ_ = opt_bsparse && n_top >= 8192 ? last_occ : 1.0;
        auto decide = [&](uint32_t n_work, bool mc, double& t_est, bool explore) -> int {      // -1: no solve for dense combinations; t_est: modelled seconds
            const uint32_t nloc = (n_work + world - 1) / world;
            const uint64_t mt_ = (nloc + sc_bm - 1) / sc_bm;
            use_schur = opt_schur && nloc && L.n_nonpiv && n_top >= (uint32_t)std::max<int64_t>(1, opt_schur_min_top) &&
                        (double)nl * (mt_ * sc_bm + nt_ * sc_bn) * kp <= (double)opt_schur_max_gb * 1e9 && mt_ * nt_ < (1ull << 31);
            bool panel_fits = false;
            if (use_schur && n_top > PANEL_T && n_top <= (uint64_t)opt_panel_max_top && (!mc || PANEL_T <= (uint32_t)BLK_TMAX)) {
                const uint64_t ntp_ = (n_top + PANEL_T - 1) / PANEL_T, kt_ = ntp_ * PANEL_T;
                // dense: every At_t at once; block-sparse: one pivot tile's occupied blocks at a time
                const double at_b = (occ_ < 1.0 || (opt_bsparse && n_top >= 8192) ? (double)nl * PANEL_T * (double)kt_ * occ_ : (double)nl * PANEL_T * PANEL_T * (double)(ntp_ * (ntp_ - 1) / 2))
                                    + 3.25 * nl * (double)kt_ * PANEL_T;
                panel_fits = at_b <= (double)opt_panel_max_gb * 1e9 &&
                             at_b + (double)nl * (mt_ * sc_bm + nt_ * sc_bn) * kt_ + (double)mt_ * sc_bm * PANEL_T * (4 + nl) <= (double)opt_schur_max_gb * 1e9;
            }
            const bool block_ok = use_schur && opt_block >= 1;
            int mode = mc ? -1 : 0;
            double est[3];
            est[0] = mc ? 1e30 : calib[0] * hitw * model_sparse(nloc, last_density, n_top, last_a, f.small != 0);
            est[1] = block_ok && (n_top >= 8192 || opt_block >= 2) ? calib[1] * model_block(nloc, n_top, have_history ? last_a : 100.0, f.small != 0) : 1e30;
            est[2] = panel_fits && opt_panel >= 1 ? calib[2] * model_panel((double)(mt_ * sc_bm), n_top, nl, PANEL_T, occ_) : 1e30;
            t_est = 1e30;
            if (opt_panel >= 2 && panel_fits) mode = 2;
            else if (opt_block >= 2 && block_ok) mode = 1;
            else if (use_schur && (have_history || mc)) {
                if (mc) {
                    // dense rows: the blocked sparse kernel finds no empty window to skip and pays one shared-memory RED per entry of A
                    // and combination, the products skip the empty blocks of A (kat15-15, 440 k pivot rows, 1480 combinations:
                    // 324 ms against ~100 ms; 458 k pivot rows, 120 combinations: 518 ms against 166 ms)
                    mode = est[2] < 1e30 ? 2 : 1;
                    if (est[mode] >= 1e30) mode = -1;
                }
                else {
                    mode = est[1] < est[0] ? 1 : 0;
                    if (est[2] < est[mode]) mode = 2;
                }
                if (explore && mode >= 0 && ++since_explore >= 8) {           // the runner-up, if it is within 1.5x
                    since_explore = 0;
                    int second = -1;
                    for (int m = 0; m < 3; ++m) if (m != mode && est[m] < 1.5 * est[mode] && (second < 0 || est[m] < est[second])) second = m;
                    if (second >= 0) mode = second;
                }
            }
            if (mode >= 0 && (have_history || mode > 0)) t_est = est[mode];
            return mode;
        };
        // Monte-Carlo combinations: K = predicted rank + margin of them instead of the n_bot rows, when that halves the rows at least
        mc_plan = 0;
        // mc = 1: where it was measured to pay (tools/mc_stats.py, kat15-15 without the pair cap: 2.4-2.8x on the steps with 440-460 k pivot
        // rows, nothing at 93 k; cyclic9-15, <= 50 k pivot rows and <= 2000 rows: the solve is launch-bound and does not get faster with fewer rows):
        // many pivot rows, half of the rows as combinations at most
        const bool mc_auto = n_top >= (uint64_t)std::max<int64_t>(1, opt_mc_min_top) && n_bot >= (uint64_t)std::max<int64_t>(256, opt_mc_min_rows);
        if (opt_mc && !(step_flags & GAMBA_CUDA_STEP_FINAL) && n_bot >= 256 && L.n_nonpiv && (opt_mc >= 2 || mc_auto)) {
            const int64_t pred = opt_mc_k > 0 ? opt_mc_k : last_rank >= 0 ? last_rank + 32 + last_rank / 8 : opt_mc >= 2 ? 64 : -1;
            if (pred > 0) {
                const uint64_t k = ((uint64_t)pred + 7) / 8 * 8;
                if (k * 2 <= n_bot) mc_plan = (uint32_t)k;
            }
        }
        // The choice between combinations and rows must be the same on every rank (the all-gather sizes follow from it): it uses
        // sizes, options and the previous step's rank only - never the calibration factors, which are rank-local timings.
        double t_mc = 1e30, t_exact = 1e30;
        int mode = mc_plan ? decide(mc_plan, true, t_mc, false) : -1;      // no runner-up runs on combinations: the sparse solve loses 150-280 ms there
        if (mode >= 0) {
            // a draw that misses the margin all the way falls back to the rows themselves ON THIS LAYOUT: their multiplier and
            // operand planes must fit the same budget, or the step is planned for the rows from the start
            const uint64_t mt_x = ((uint64_t)(n_bot + world - 1) / world + sc_bm - 1) / sc_bm;
            if ((double)nl * (mt_x * sc_bm + nt_ * sc_bn) * ((double)kp + PANEL_T) > (double)opt_schur_max_gb * 1e9) mode = -1;
        }
        if (mode < 0) { mc_plan = 0; mode = decide(n_bot, false, t_exact, true); }
        use_block = mode == 1; use_panel = mode == 2;
        if (use_block) tmax = BLK_TMAX;
        if (use_panel) tmax = PANEL_T;
    }
    if (opt_tile_cols > 0 && !use_panel) tmax = std::min<uint32_t>(tmax, std::max<uint32_t>(gran, (uint32_t)opt_tile_cols / gran * gran));
    auto up = [&](uint32_t x) { return std::max(gran, (x + gran - 1) / gran * gran); };
    // pivot tiles of equal width; the non-pivot range is cut with the same width
    uint32_t ntp = n_top ? (n_top + tmax - 1) / tmax : 0;
    uint32_t t = ntp ? up((n_top + ntp - 1) / ntp) : gran;
    if (L.n_nonpiv > t) {
        uint32_t ntn = (L.n_nonpiv + tmax - 1) / tmax;
        t = std::max(t, up((L.n_nonpiv + ntn - 1) / ntn));
    }
    t = std::min(t, tmax);
    if (use_panel) t = PANEL_T;          // the products address the operands by tile: fixed width, multiple of 128
    L.tile_cols = t;
    L.n_tiles_piv = (n_top + t - 1) / t;
    L.n_tiles_nonpiv = (L.n_nonpiv + t - 1) / t;
    L.n_tiles = std::max(1u, L.n_tiles_piv + L.n_tiles_nonpiv);
    L.n_windows = (n_top + WINDOW - 1) / WINDOW;
    ctr.tile_cols = t; ctr.n_tiles = L.n_tiles; ctr.acc_bits = 32 * words_per_col;
}

// Sizes of a step; with world_size > 1 and a broadcast callback only rank 0 holds the step on the host
// and this header travels first ("ncclBroadcast of the flattened A|B").
struct StepHdr { uint32_t n_top, n_bot, n_cols, n_arr; uint64_t nnz, pool_len, n_src; uint32_t flags, col_map_len, has_row_src, pad; };

void gamba_cuda_engine::stage(const gamba_cuda_step_in* in) {
    const double t0 = wall_s();
    staged = false; eliminated = false;
    const bool use_bcast = cfg.world_size > 1 && bcast_fn != nullptr;
    const bool have = !use_bcast || cfg.rank == 0;       // this rank uploads the step from its host buffers
    if (have && !in) throw std::runtime_error("no step passed (only ranks > 0 of a broadcasting group may pass NULL)");
    StepHdr hd{};
    // GAMBA_CUDA_STEP_DEVICE_INDICES: the index arrays of the step already live on this device and are used where they are
    const bool dev_idx = have && (in->flags & GAMBA_CUDA_STEP_DEVICE_INDICES) != 0;
    if (dev_idx && use_bcast) throw std::runtime_error("device-resident index arrays cannot be combined with the broadcast callback");
    // GAMBA_CUDA_STEP_DEVICE_COLIDX: col_idx alone is device memory (the handle array of gamba_cuda_rowgen_add_rows)
    const bool dev_col = have && !dev_idx && (in->flags & GAMBA_CUDA_STEP_DEVICE_COLIDX) != 0;
    const bool alias_col = dev_idx || (dev_col && !use_bcast);
    if (have) {
        if (in->n_cols < in->n_top) throw std::runtime_error("n_cols < n_top");
        if (!dev_idx && in->row_ptr[in->n_top + in->n_bot] != in->nnz) throw std::runtime_error("row_ptr[n_rows] != nnz");
        hd.n_top = in->n_top; hd.n_bot = in->n_bot; hd.n_cols = in->n_cols; hd.n_arr = in->n_coeff_arrays;
        hd.nnz = in->nnz; hd.flags = in->flags;
        hd.n_src = in->col_idx_len ? in->col_idx_len : in->nnz;
        hd.col_map_len = in->col_map ? in->col_map_len : 0; hd.has_row_src = in->row_src != nullptr;
        if (hd.n_src < hd.nnz && !hd.has_row_src) throw std::runtime_error("col_idx_len < nnz");
        for (uint32_t i = 0; i < in->n_coeff_arrays; ++i) hd.pool_len += in->coeff_len[i];
    }
    if (use_bcast) {
        d_hdr.ensure(sizeof(StepHdr)); h_hdr.ensure(sizeof(StepHdr));
        if (have) {
            std::memcpy(h_hdr.p, &hd, sizeof hd);
            GCU_CHECK(cudaMemcpyAsync(d_hdr.p, h_hdr.p, sizeof hd, cudaMemcpyHostToDevice, stream));
        }
        if (bcast_fn(bcast_user, d_hdr.p, sizeof hd, 0, stream)) throw std::runtime_error("broadcast of the step header failed");
        GCU_CHECK(cudaMemcpyAsync(h_hdr.p, d_hdr.p, sizeof hd, cudaMemcpyDeviceToHost, stream));
        GCU_CHECK(cudaStreamSynchronize(stream));
        std::memcpy(&hd, h_hdr.p, sizeof hd);
    }
    const uint32_t n_rows = hd.n_top + hd.n_bot;
    choose_layout(hd.n_top, hd.n_bot, hd.n_cols, hd.flags, hd.nnz);
    flags = hd.flags & ~(GAMBA_CUDA_STEP_DEVICE_INDICES | GAMBA_CUDA_STEP_DEVICE_COLIDX); nnz = hd.nnz;
    if ((uint64_t)n_rows * L.n_tiles + 1 >= (1ull << 32)) throw std::runtime_error("segment table too large");
    if (nnz + 1ull * n_rows * L.n_tiles + 4 >= (1ull << 32)) throw std::runtime_error("step has too many non-zeros for 32-bit entry offsets");
    if ((uint64_t)8 * L.n_tiles * 4 > 160 * 1024) throw std::runtime_error("too many column tiles for the staging kernels");

    double tq = wall_s(); dbg_[0] += tq - t0;
    // shared coefficient arrays -> one u32 pool. With the option "coeff_cache" the pool is an append-only
    // device arena keyed by (host pointer, length): the caller promises that coefficient arrays are immutable
    // while the engine lives (source/matrix.hpp:189-193: they alias basis storage), so an array crosses
    // PCIe once per run instead of once per step.
    const uint32_t na = hd.n_arr;
    const bool cached = opt_coeff_cache && !use_bcast;
    h_off.ensure((na + 1) * sizeof(uint64_t));
    uint64_t* off = h_off.as<uint64_t>();
    uint64_t tot = hd.pool_len;          // entries to upload this step
    uint64_t pool_base = 0;              // where they go in d_pool
    std::vector<uint32_t> fresh;         // arrays not yet on the device
    if (cached) {
        tot = 0; pool_base = cache_used;
        fresh.reserve(na);
        for (uint32_t i = 0; i < na; ++i) {
            auto it = cache.find(in->coeff_ptr[i]);
            const uint32_t n_i = in->coeff_len[i];
            auto word = [&](uint32_t k) -> uint32_t {
                return cfg.coeff_bytes == 4 ? ((const uint32_t*)in->coeff_ptr[i])[k] : cfg.coeff_bytes == 2 ? ((const uint16_t*)in->coeff_ptr[i])[k]
                                                                                                          : ((const uint8_t*)in->coeff_ptr[i])[k];
            };
            const uint32_t w0 = n_i ? word(0) : 0u, w1 = n_i ? word(n_i - 1) : 0u;
            if (it != cache.end() && it->second.len == n_i && it->second.first == w0 && it->second.last == w1) { off[i] = it->second.off; continue; }
            off[i] = cache_used + tot; tot += n_i;
            cache[in->coeff_ptr[i]] = CacheEntry{off[i], n_i, w0, w1};
            fresh.push_back(i);
        }
        off[na] = cache_used + tot;
        if ((cache_used + tot) * sizeof(uint32_t) > d_pool.cap && d_pool.va) d_pool.ensure((cache_used + tot) * 2 * sizeof(uint32_t));   // grows in place
        if ((cache_used + tot) * sizeof(uint32_t) > d_pool.cap) {      // grow the arena, keeping what is there
            gcu::DevBuf bigger;
            bigger.ensure((cache_used + tot) * 2 * sizeof(uint32_t));
            if (cache_used) GCU_CHECK(cudaMemcpyAsync(bigger.p, d_pool.p, cache_used * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
            GCU_CHECK(cudaStreamSynchronize(stream));
            std::swap(bigger.p, d_pool.p); std::swap(bigger.cap, d_pool.cap);
            std::swap(bigger.va, d_pool.va); std::swap(bigger.va_size, d_pool.va_size); std::swap(bigger.chunks, d_pool.chunks);
        }
        cache_used += tot;
    } else if (have) {
        uint64_t t = 0;
        for (uint32_t i = 0; i < na; ++i) { off[i] = t; t += in->coeff_len[i]; }
        off[na] = t;
    }
    h_pool.ensure(std::max<uint64_t>(tot, 1) * sizeof(uint32_t));
    uint32_t* pool = h_pool.as<uint32_t>();
    auto pack_one = [&](uint32_t i, uint32_t* dst) {
        const uint32_t n = in->coeff_len[i];
        switch (cfg.coeff_bytes) {
            case 4: std::memcpy(dst, in->coeff_ptr[i], (size_t)n * 4); break;
            case 2: { const uint16_t* s = (const uint16_t*)in->coeff_ptr[i]; for (uint32_t k = 0; k < n; ++k) dst[k] = s[k]; } break;
            default: { const uint8_t* s = (const uint8_t*)in->coeff_ptr[i]; for (uint32_t k = 0; k < n; ++k) dst[k] = s[k]; } break;
        }
    };
    auto pack_pool = [&]() {
        if (cached) { uint64_t t = 0; for (uint32_t i : fresh) { pack_one(i, pool + t); t += in->coeff_len[i]; } }
        else if (have) for (uint32_t i = 0; i < na; ++i) pack_one(i, pool + off[i]);
    };
    if (!dev_idx) {
        d_row_ptr.ensure((n_rows + 1) * sizeof(uint64_t));
        if (!alias_col) d_col.ensure(std::max<uint64_t>(hd.n_src, 1) * sizeof(uint32_t));
        if (hd.has_row_src) d_row_src.ensure(std::max<uint32_t>(n_rows, 1) * sizeof(uint64_t));
        if (hd.col_map_len) d_col_map.ensure((size_t)hd.col_map_len * sizeof(uint32_t));
        d_coeff_id.ensure(std::max<uint32_t>(n_rows, 1) * sizeof(uint32_t));
    }
    d_coeff_off.ensure((na + 1) * sizeof(uint64_t));
    if (!cached) d_pool.ensure(std::max<uint64_t>(tot, 1) * sizeof(uint32_t));
    h2d_bytes = 0;
    dbg_[1] += wall_s() - tq; tq = wall_s();
    if (have) {
        if (!dev_idx) {
            GCU_CHECK(cudaMemcpyAsync(d_row_ptr.p, in->row_ptr, (n_rows + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
            if (!alias_col) GCU_CHECK(cudaMemcpyAsync(d_col.p, in->col_idx, hd.n_src * sizeof(uint32_t), dev_col ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
            if (hd.has_row_src) GCU_CHECK(cudaMemcpyAsync(d_row_src.p, in->row_src, n_rows * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
            if (hd.col_map_len) GCU_CHECK(cudaMemcpyAsync(d_col_map.p, in->col_map, (size_t)hd.col_map_len * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
            GCU_CHECK(cudaMemcpyAsync(d_coeff_id.p, in->coeff_id, n_rows * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        }
        // the coefficient pool is packed while the column indices (page-locked at the caller: a plain DMA) are in flight
        pack_pool();
        GCU_CHECK(cudaMemcpyAsync(d_coeff_off.p, off, (na + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
        if (tot) GCU_CHECK(cudaMemcpyAsync(d_pool.as<uint32_t>() + pool_base, pool, tot * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        h2d_bytes = (na + 1) * 8ull + tot * 4ull;
        if (!dev_idx) h2d_bytes += (n_rows + 1) * 8ull + (dev_col ? 0ull : hd.n_src * 4ull) + n_rows * 4ull + (hd.has_row_src ? n_rows * 8ull : 0) + hd.col_map_len * 4ull;
    }
    if (use_bcast) {   // A|B and C|D replicated over NVLink instead of one H2D copy per rank
        const std::pair<void*, size_t> bufs[] = {{d_row_ptr.p, (n_rows + 1) * sizeof(uint64_t)}, {d_col.p, hd.n_src * sizeof(uint32_t)},
                                                 {d_row_src.p, hd.has_row_src ? n_rows * sizeof(uint64_t) : 0}, {d_col_map.p, (size_t)hd.col_map_len * sizeof(uint32_t)},
                                                 {d_coeff_id.p, n_rows * sizeof(uint32_t)}, {d_coeff_off.p, (na + 1) * sizeof(uint64_t)},
                                                 {d_pool.p, tot * sizeof(uint32_t)}};
        for (auto const& b : bufs)
            if (b.second && bcast_fn(bcast_user, b.first, b.second, 0, stream)) throw std::runtime_error("broadcast of the step failed");
    }

    dbg_[2] += wall_s() - tq; tq = wall_s();
    const uint64_t n_cnt = (uint64_t)n_rows * L.n_tiles + 1;
    const uint64_t n_diag = std::max<uint64_t>((uint64_t)L.n_windows * WINDOW * WINDOW, 1);
    d_seg.ensure(n_cnt * sizeof(uint32_t));
    d_ent.ensure((nnz + 1ull * n_rows * L.n_tiles + 4) * sizeof(uint2));   // segments padded to an even number of entries
    d_diag.ensure(n_diag * sizeof(uint32_t));
    d_dinv.ensure(n_diag * sizeof(uint32_t));
    d_dnz.ensure(std::max<uint32_t>(L.n_windows, 1) * sizeof(uint32_t));
    d_first.ensure((std::max<uint32_t>(L.n_bot, 1) + 1) * sizeof(uint32_t));
    dbg_[3] += wall_s() - tq; tq = wall_s();
    GCU_CHECK(cudaEventRecord(ev[0], stream));
    GCU_CHECK(cudaMemsetAsync(d_seg.p, 0, n_cnt * sizeof(uint32_t), stream));
    GCU_CHECK(cudaMemsetAsync(d_diag.p, 0, n_diag * sizeof(uint32_t), stream));

    PrepArgs pa{};
    pa.L = L; pa.f = f;
    pa.row_ptr = dev_idx ? in->row_ptr : d_row_ptr.as<uint64_t>(); pa.col = alias_col ? in->col_idx : d_col.as<uint32_t>();
    pa.coeff_id = dev_idx ? in->coeff_id : d_coeff_id.as<uint32_t>();
    s_row_ptr = pa.row_ptr;
    pa.coeff_off = d_coeff_off.as<uint64_t>(); pa.pool = d_pool.as<uint32_t>();
    pa.seg = d_seg.as<uint32_t>(); pa.ent = d_ent.as<uint2>();
    pa.diag = d_diag.as<uint32_t>(); pa.dinv = d_dinv.as<uint32_t>(); pa.dnz = d_dnz.as<uint32_t>();
    pa.row_src = hd.has_row_src ? (dev_idx ? in->row_src : d_row_src.as<uint64_t>()) : nullptr;
    pa.col_map = hd.col_map_len ? (dev_idx ? in->col_map : d_col_map.as<uint32_t>()) : nullptr;
    pa.first_piv = d_first.as<uint32_t>(); pa.first_min = d_first.as<uint32_t>() + std::max<uint32_t>(L.n_bot, 1);
    GCU_CHECK(cudaMemsetAsync(pa.first_min, 0xff, sizeof(uint32_t), stream));
    if (n_rows) {
        const size_t sh = (size_t)8 * L.n_tiles * sizeof(uint32_t);
        if (sh > 48 * 1024) {
            GCU_CHECK(cudaFuncSetAttribute(prep_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
            GCU_CHECK(cudaFuncSetAttribute(prep_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        }
        const uint32_t nblk = (n_rows + 7) / 8;
        prep_count_kernel<<<nblk, 256, sh, stream>>>(pa);
        GCU_CHECK(cudaGetLastError());
        size_t tmp_bytes = 0;
        GCU_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, pa.seg, pa.seg, (int64_t)n_cnt, stream));
        d_scan_tmp.ensure(tmp_bytes);
        GCU_CHECK(cub::DeviceScan::ExclusiveSum(d_scan_tmp.p, tmp_bytes, pa.seg, pa.seg, (int64_t)n_cnt, stream));
        prep_fill_kernel<<<nblk, 256, sh, stream>>>(pa);
        GCU_CHECK(cudaGetLastError());
        ctr.kernel_launches += 3;
        if (L.n_windows) {
            prep_diag_inverse_kernel<<<(L.n_windows + 7) / 8, 256, 0, stream>>>(pa);
            GCU_CHECK(cudaGetLastError());
            ctr.kernel_launches += 1;
        }
    }
    // nnz of A (for the next step's choice between the sparse kernels and the panel products)
    h_misc.ensure(64);
    d_counts.ensure(256);
    if (n_rows && L.n_top) {
        prep_nnz_a_kernel<<<1, 32, 0, stream>>>(pa, reinterpret_cast<unsigned long long*>(d_counts.as<char>() + 128));
        GCU_CHECK(cudaGetLastError());
        GCU_CHECK(cudaMemcpyAsync(h_misc.as<char>() + 48, d_counts.as<char>() + 128, 8, cudaMemcpyDeviceToHost, stream));
    }
    GCU_CHECK(cudaEventRecord(ev[1], stream));
    GCU_CHECK(cudaStreamSynchronize(stream));
    if (n_rows && L.n_top) { cur_a = (double)*reinterpret_cast<unsigned long long*>(h_misc.as<char>() + 48) / L.n_top; }
    float ms = 0;
    GCU_CHECK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
    t_prep_dev = ms * 1e-3;
    dbg_[4] += wall_s() - tq;
    t_stage_wall = wall_s() - t0;
    ctr.seconds_stage_last = t_stage_wall;
    staged = true;
}

// One launch of the limb-split tensor-core product (kernels_schur_tc.cuh)
template <int EPI>
static void launch_limb_gemm(bool small, const CUtensorMap& ma, const CUtensorMap& mb, const gcu::TcArgs& ta, uint32_t batches, cudaStream_t st) {
    using namespace gcu;
    const dim3 grid(ta.m_tiles * ta.n_tiles, std::max(1u, batches));
    if (small) {
        GCU_CHECK(cudaFuncSetAttribute(limb_gemm_tc_kernel<2, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<2>()));
        limb_gemm_tc_kernel<2, EPI><<<grid, TC_THREADS, tc_smem_bytes<2>(), st>>>(ma, mb, ta);
    } else {
        GCU_CHECK(cudaFuncSetAttribute(limb_gemm_tc_kernel<4, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<4>()));
        limb_gemm_tc_kernel<4, EPI><<<grid, TC_THREADS, tc_smem_bytes<4>(), st>>>(ma, mb, ta);
    }
    GCU_CHECK(cudaGetLastError());
}

void gamba_cuda_engine::eliminate() {
    if (!staged) throw std::runtime_error("no staged step");
    const uint32_t world = std::max(1u, cfg.world_size), rank = cfg.rank;
    const uint32_t n_work = mc_k ? mc_k : L.n_bot;       // rows of C|D, or random combinations of them
    n_local_max = (n_work + world - 1) / world;
    n_local = n_work > rank ? (n_work - rank + world - 1) / world : 0;
    n_gather_rows = mc_base + (world > 1 ? n_local_max * world : n_work);
    if (mc_k && !(use_schur && (use_block || use_panel) && n_local)) throw std::runtime_error("Monte-Carlo combinations need one of the blocked solves");

    const uint32_t words_per_col = f.small ? 1u : 2u;
    const size_t acc_words = (((size_t)L.tile_cols + 1) * words_per_col + 3) & ~(size_t)3;
    const size_t smem = (acc_words + (size_t)elim_warps * ELIM_SCRATCH_WORDS) * sizeof(uint32_t);
    const int elim_threads = elim_warps * 32;
    auto kern = elim_warps == 32 ? (f.small ? elim_kernel<true, 32> : elim_kernel<false, 32>)
              : elim_warps == 2 ? (f.small ? elim_kernel<true, 2> : elim_kernel<false, 2>)
              : elim_warps == 4 ? (f.small ? elim_kernel<true, 4> : elim_kernel<false, 4>)
                                : (f.small ? elim_kernel<true, 8> : elim_kernel<false, 8>);
    const uint32_t nl = f.small ? 2u : 4u;
    const uint32_t bn = f.small ? TcCfg<2>::BN : TcCfg<4>::BN;
    const uint32_t tkc = f.small ? TcCfg<2>::KCHUNK : TcCfg<4>::KCHUNK;
    const uint32_t kchunk = opt_schur_kchunk > 0 ? (uint32_t)std::min<int64_t>(opt_schur_kchunk, tkc) : tkc;
    const bool schur = use_schur && n_local;
    bool block = schur && use_block;
    bool panel = schur && use_panel;
    elim_grid = 1;
    if (!panel) {
        GCU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        GCU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, elim_threads, smem));
        if (occ < 1) throw std::runtime_error("elimination kernel does not fit on an SM");
        if (opt_ctas_per_sm > 0) occ = std::min<int>(occ, (int)opt_ctas_per_sm);
        elim_grid = std::max(1, std::min<int>(prop.multiProcessorCount * occ, (int)std::max(1u, n_local)));
    }
    const uint64_t n_slots = (uint64_t)elim_grid;

    const uint32_t cap_rows = std::max(n_gather_rows, mc_k ? mc_cap_rows : 0u);     // a later draw must find the earlier combinations where they are
    d_dprime.ensure(std::max<uint64_t>((uint64_t)cap_rows * L.ld_dprime, 1) * sizeof(uint32_t));
    d_row_nz.ensure(std::max<uint32_t>(cap_rows, 1) * sizeof(uint32_t));
    if (!(schur && (block || panel))) {     // the blocked kernel keeps its multipliers in d_mt, the blocked solve needs no lists
        d_list_k.ensure(std::max<uint64_t>(n_slots * L.n_top, 1) * sizeof(uint32_t));
        d_list_m.ensure(std::max<uint64_t>(n_slots * L.n_top, 1) * sizeof(uint32_t));
    }
    d_queue_stats.ensure(64);
    GCU_CHECK(cudaMemsetAsync(d_queue_stats.p, 0, 64, stream));
    if (n_gather_rows > mc_base) GCU_CHECK(cudaMemsetAsync(d_row_nz.as<uint32_t>() + mc_base, 0, (size_t)(n_gather_rows - mc_base) * sizeof(uint32_t), stream));

    // local block of D' sits at its final place inside the gather buffer (behind the combinations of earlier draws)
    const uint64_t loc_row0 = (uint64_t)mc_base + (world > 1 ? (uint64_t)rank * n_local_max : 0);
    uint32_t* dloc = d_dprime.as<uint32_t>() + loc_row0 * L.ld_dprime;
    uint32_t* nzloc = d_row_nz.as<uint32_t>() + loc_row0;

    ElimArgs ea{};
    ea.L = L; ea.f = f;
    ea.mu32 = (uint32_t)((1ull << 32) / f.p);
    ea.negp = 0u - f.p;
    // every add is < 2p (p < 2^16, lazy Barrett) or < 2^16 (lo/hi halves): normalise before 2^32 is reached
    ea.norm_every = f.small ? (uint32_t)std::min<uint64_t>(0xffffffffull / (2ull * f.p) - 8 * 32 - 2, 1u << 30) : 65536u - 8 * 32 - 2;
    ea.seg = d_seg.as<uint32_t>(); ea.ent = d_ent.as<uint2>(); ea.diag = d_dinv.as<uint32_t>();
    ea.dnz = d_dnz.as<uint32_t>();
    ea.first_piv = d_first.as<uint32_t>();
    ea.dprime = dloc; ea.row_nz = nzloc;
    ea.list_k = d_list_k.as<uint32_t>(); ea.list_m = d_list_m.as<uint32_t>();
    ea.queue = d_queue_stats.as<uint32_t>();
    ea.stats = reinterpret_cast<unsigned long long*>(d_queue_stats.as<char>() + 16);
    ea.row_ptr = s_row_ptr;
    uint32_t mc_first = 0;           // smallest pivot column any row to reduce holds: where every combination starts
    if (mc_k) { GCU_CHECK(cudaMemcpyAsync(&mc_first, d_first.as<uint32_t>() + std::max<uint32_t>(L.n_bot, 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)); GCU_CHECK(cudaStreamSynchronize(stream)); }
    ea.row_begin = rank; ea.row_step = world; ea.n_local = n_local;
    ea.t_lo = 0; ea.t_hi = L.n_tiles;

    // Schur mode: the B half of the elimination is a dense product on the tensor cores (kernels_schur_tc.cuh).
    // k pitch of the limb planes: n_top rounded up to 128 bytes, or to whole pivot tiles for the blocked solve
    const uint32_t T = PANEL_T, ntp = L.n_tiles_piv;
    const uint32_t kt = ntp * T;                                    // blocked solve: pivot range padded to whole tiles
    const uint32_t kp = panel ? kt : (L.n_top + TC_BK - 1) / TC_BK * TC_BK;
    const uint32_t k_used = (L.n_top + TC_BK - 1) / TC_BK * TC_BK;   // contraction depth of the Schur product
    const uint32_t mp = (n_local + 127u) / 128u * 128u, np = (L.n_nonpiv + 127u) / 128u * 128u;   // rows per limb plane
    const uint64_t mx_plane = (uint64_t)mp * kp, bt_plane = (uint64_t)np * kp;
    const uint64_t tri_plane = (uint64_t)kt * T;
    GCU_CHECK(cudaEventRecord(ev[2], stream));
    const uint64_t mc_ldc = (uint64_t)ntp * L.tile_cols;
    if (mc_k) {
        // the combinations as dense rows: C half into d_mc_c, D half straight into D' (kernels_mc.cuh)
        const uint32_t R = f.small ? BlkRows<true>::R : BlkRows<false>::R;
        McRowsArgs ma{};
        ma.L = L; ma.f = f; ma.mu32 = ea.mu32; ma.negp = ea.negp; ma.norm_every = ea.norm_every;
        ma.seg = ea.seg; ma.ent = ea.ent;
        ma.nbp = (L.n_bot + 31u) / 32u * 32u; ma.n_q = n_local; ma.n_q_pad = (n_local + R - 1) / R * R;
        ma.q_global0 = mc_base + rank * n_local_max;
        ma.seed = cfg.seed * 0x9e3779b97f4a7c15ull + 0xd1342543de82ef95ull * mc_draw;
        d_mc_sel.ensure((uint64_t)ma.n_q_pad * ma.nbp * sizeof(uint32_t));
        d_mc_c.ensure(std::max<uint64_t>((uint64_t)n_local * mc_ldc, 1) * sizeof(uint32_t));
        ma.s = d_mc_sel.as<uint32_t>(); ma.dst_c = d_mc_c.as<uint32_t>(); ma.ldc = mc_ldc; ma.dst_d = dloc; ma.ldd = L.ld_dprime;
        GCU_CHECK(cudaMemsetAsync(dloc, 0, (uint64_t)n_local * L.ld_dprime * sizeof(uint32_t), stream));
        mc_gen_kernel<<<(uint32_t)(((uint64_t)ma.n_q_pad * ma.nbp + 255) / 256), 256, 0, stream>>>(ma);
        GCU_CHECK(cudaGetLastError());
        auto mk = f.small ? mc_rows_kernel<true> : mc_rows_kernel<false>;
        const size_t msm = (size_t)R * BLK_TS * (f.small ? 1 : 2) * 4;
        GCU_CHECK(cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
        if (L.tile_cols > (uint32_t)BLK_TMAX) throw std::runtime_error("Monte-Carlo combinations: column tile too wide");
        mk<<<dim3(ma.n_q_pad / R, L.n_tiles), BLK_THREADS, msm, stream>>>(ma);
        GCU_CHECK(cudaGetLastError());
        ctr.kernel_launches += 2;
    }
    // Block-sparse operands (kernels_schur.cuh, BsArgs): which blocks of 128 pivot rows x bn columns of B (and, for the blocked
    // solve, of A right of a row's own tile) hold an entry at all. The products then skip the empty ones.
    const bool bs = schur && opt_bsparse && L.tile_cols % bn == 0 && L.n_top >= 8192;
    const uint32_t per_tile = bs ? L.tile_cols / bn : 0;
    // the blocks of A are marked in every mode (one pass over A's entries): their occupied fraction is what the next step's
    // choice of the solve needs
    const uint32_t n_cb_piv = bs ? ntp * per_tile : 0, n_cb = n_cb_piv + (bs ? L.n_tiles_nonpiv * per_tile : 0);
    const uint32_t kb_words = ((L.n_top + 127u) / 128u + 31u) / 32u;
    uint64_t nb_total = 0, nb_a = 0, bs_tile_max = 0;
    BsArgs bsa{};
    if (bs) {
        d_bs_bitmap.ensure((uint64_t)n_cb * kb_words * 4); d_bs_rank.ensure((uint64_t)n_cb * kb_words * 4); d_bs_base.ensure((uint64_t)(n_cb + 1) * 4);
        GCU_CHECK(cudaMemsetAsync(d_bs_bitmap.p, 0, (uint64_t)n_cb * kb_words * 4, stream));
        GCU_CHECK(cudaMemsetAsync(d_bs_base.p, 0, (uint64_t)(n_cb + 1) * 4, stream));
        bsa.L = L; bsa.seg = d_seg.as<uint32_t>(); bsa.ent = d_ent.as<uint2>(); bsa.cbw = bn; bsa.kb_words = kb_words;
        bsa.n_cb_piv = n_cb_piv; bsa.n_cb = n_cb; bsa.bitmap = d_bs_bitmap.as<uint32_t>(); bsa.rank = d_bs_rank.as<uint32_t>();
        bsa.blk_base = d_bs_base.as<uint32_t>(); bsa.with_a = 1u;
        bs_mark_kernel<<<(L.n_top + 7) / 8, 256, 0, stream>>>(bsa);
        GCU_CHECK(cudaGetLastError());
        bs_rank_kernel<<<(n_cb + 7) / 8, 256, 0, stream>>>(bsa);
        GCU_CHECK(cudaGetLastError());
        size_t tmp_bytes = 0;
        GCU_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, bsa.blk_base, bsa.blk_base, (int64_t)(n_cb + 1), stream));
        d_scan_tmp.ensure(tmp_bytes);
        GCU_CHECK(cub::DeviceScan::ExclusiveSum(d_scan_tmp.p, tmp_bytes, bsa.blk_base, bsa.blk_base, (int64_t)(n_cb + 1), stream));
        uint32_t cnt[2] = {0, 0};
        bs_tile_base.assign(ntp + 1, 0);
        GCU_CHECK(cudaMemcpyAsync(&cnt[0], bsa.blk_base + n_cb, 4, cudaMemcpyDeviceToHost, stream));
        GCU_CHECK(cudaMemcpyAsync(&cnt[1], bsa.blk_base + n_cb_piv, 4, cudaMemcpyDeviceToHost, stream));
        if (panel)      // first block of every pivot tile: the operand of the earlier-tile product is built one tile at a time
            GCU_CHECK(cudaMemcpy2DAsync(bs_tile_base.data(), 4, bsa.blk_base, (size_t)per_tile * 4, 4, ntp + 1, cudaMemcpyDeviceToHost, stream));
        GCU_CHECK(cudaStreamSynchronize(stream));
        nb_total = cnt[0]; nb_a = cnt[1];
        for (uint32_t t = 0; t < ntp && panel; ++t) bs_tile_max = std::max<uint64_t>(bs_tile_max, bs_tile_base[t + 1] - bs_tile_base[t]);
        ctr.kernel_launches += 3;
        const double tri_cells = (double)n_cb_piv * ((L.n_top + 127u) / 128u) * 0.5, b_cells = (double)(n_cb - n_cb_piv) * ((L.n_top + 127u) / 128u);
        last_occ = tri_cells > 0 ? (double)nb_a / tri_cells : (b_cells > 0 ? (double)(nb_total - nb_a) / b_cells : 1.0);
        last_occ = std::min(1.0, std::max(0.01, last_occ));
        // blocked solve whose operands do not fit after all: the sparse kernel for blocks of rows on the same layout
        if (panel && (double)nl * (nb_total - nb_a + bs_tile_max) * bn * 128.0 > (double)opt_schur_max_gb * 1e9) { panel = false; block = true; use_panel = false; use_block = true; }
    }
    const uint64_t bs_plane = (nb_total - nb_a) * bn * 128ull;          // B alone: the blocks of A are built one pivot tile at a time
    const uint64_t bs_at_plane = std::max<uint64_t>(bs_tile_max, 1) * bn * 128ull;
    OpArgs oa_at{};
    void (*opk)(OpArgs) = nullptr;
    size_t op_sm = 0;
    uint32_t chunks = 0;
    // buffers: growing one may go to the driver
    if (schur) {
        d_mx.ensure(nl * mx_plane);
        if (bs) { d_bs_op.ensure(std::max<uint64_t>(nl * bs_plane, 16)); d_bs_klist.ensure(std::max<uint64_t>(nb_total, 1) * 4); if (panel) d_bs_at.ensure(nl * bs_at_plane); }
        else d_bt.ensure(nl * bt_plane);
        if (panel) {     // operands of the products with the earlier tiles: tile t takes [nl][T][t * T] bytes
            if (!bs) {
                at_off.assign(ntp + 1, 0);
                for (uint32_t t = 1; t <= ntp; ++t) at_off[t] = at_off[t - 1] + (t > 1 ? (uint64_t)nl * T * (uint64_t)(t - 1) * T : 0);
                at_bytes = at_off[ntp];
                d_at.ensure(std::max<uint64_t>(at_bytes, 16));
                d_at_off.ensure(at_off.size() * sizeof(uint64_t));
            }
            d_cd.ensure((uint64_t)mp * T * sizeof(uint32_t));
            d_cx.ensure((uint64_t)nl * mp * T);
            d_wr.ensure(nl * tri_plane); d_vt.ensure(nl * tri_plane); d_adt.ensure(nl * tri_plane);
            d_gx.ensure(std::max<uint64_t>(nl * tri_plane / 4, 16));
        }
    }
    if (schur) {
        // operands of the products, rebuilt for every reduction of the step (they are part of the algorithm, not of the
        // hand-over): operand_build_kernel writes every byte of Bt / At / Adt, so only the planes written piecemeal are zeroed
        GCU_CHECK(cudaMemsetAsync(d_mx.p, 0, nl * mx_plane, stream));
        ea.mx = d_mx.as<uint8_t>(); ea.mx_plane = mx_plane; ea.mx_kp = kp; ea.mx_nl = nl;
        OpArgs oa{};
        oa.L = L; oa.seg = d_seg.as<uint32_t>(); oa.ent = d_ent.as<uint2>();
        opk = f.small ? operand_build_kernel<2> : operand_build_kernel<4>;
        op_sm = f.small ? op_smem<2>() : op_smem<4>();
        const uint32_t och = f.small ? op_ch<2>() : op_ch<4>();
        GCU_CHECK(cudaFuncSetAttribute(opk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op_sm));
        chunks = (L.tile_cols + och - 1) / och;
        if (bs) {
            bsa.klist = d_bs_klist.as<uint32_t>();
            bs_list_kernel<<<(n_cb + 7) / 8, 256, 0, stream>>>(bsa);
            GCU_CHECK(cudaGetLastError());
            oa.bitmap = bsa.bitmap; oa.rank = bsa.rank; oa.blk_base = bsa.blk_base; oa.kb_words = kb_words; oa.cbw = bn; oa.n_cb_piv = n_cb_piv;
            oa.out = d_bs_op.as<uint8_t>(); oa.plane = bs_plane; oa.rows = 0xffffffffu; oa.pitch = 0;
            oa.mode = OP_B; oa.id0 = (uint32_t)nb_a;
            opk<<<dim3(kp / OP_KB, chunks, L.n_tiles_nonpiv), 256, op_sm, stream>>>(oa);
            GCU_CHECK(cudaGetLastError());
            ctr.kernel_launches += 2;
            oa_at = oa;                 // the blocks of A: per pivot tile, inside the tile loop
            oa_at.mode = OP_AT; oa_at.out = d_bs_at.as<uint8_t>(); oa_at.plane = bs_at_plane;
            oa.bitmap = nullptr; oa.id0 = 0;
        } else {
            oa.mode = OP_B; oa.out = d_bt.as<uint8_t>(); oa.plane = bt_plane; oa.pitch = kp; oa.rows = np;
            if ((uint64_t)L.n_tiles_nonpiv * L.tile_cols < np)        // padding rows beyond the last tile (tile width not a multiple of 128)
                for (uint32_t l = 0; l < nl; ++l)
                    GCU_CHECK(cudaMemsetAsync(d_bt.as<uint8_t>() + l * bt_plane + (uint64_t)L.n_tiles_nonpiv * L.tile_cols * kp, 0,
                                              (uint64_t)(np - L.n_tiles_nonpiv * L.tile_cols) * kp, stream));
            opk<<<dim3(kp / OP_KB, chunks, L.n_tiles_nonpiv), 256, op_sm, stream>>>(oa);
            GCU_CHECK(cudaGetLastError());
            ctr.kernel_launches += 1;
            if (panel && ntp > 1) {
                GCU_CHECK(cudaMemcpyAsync(d_at_off.p, at_off.data(), at_off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
                oa.mode = OP_AT; oa.out = d_at.as<uint8_t>(); oa.at_off = d_at_off.as<uint64_t>();
                opk<<<dim3((ntp - 1) * (T / OP_KB), chunks, ntp), 256, op_sm, stream>>>(oa);
                GCU_CHECK(cudaGetLastError());
                ctr.kernel_launches += 1;
            }
        }
        if (panel) {
            oa.mode = OP_ADT; oa.out = d_adt.as<uint8_t>(); oa.plane = tri_plane; oa.pitch = T; oa.at_off = nullptr; oa.rows = 0;
            opk<<<dim3(T / OP_KB, chunks, ntp), 256, op_sm, stream>>>(oa);
            GCU_CHECK(cudaGetLastError());
            GCU_CHECK(cudaMemsetAsync(d_wr.p, 0, nl * tri_plane, stream));
            GCU_CHECK(cudaMemsetAsync(d_vt.p, 0, nl * tri_plane, stream));
            ctr.kernel_launches += 1;
        }
    }
    GCU_CHECK(cudaEventRecord(ev[6], stream));
    if (panel) {
        // ---- blocked triangular solve on the tensor cores (kernels_tri.cuh) ----
        // 1. inverse diagonal tiles: 128 x 128 blocks on the CUDA cores, then recursive doubling as batched products
        {
            TriBaseArgs tb{};
            tb.L = L; tb.f = f; tb.seg = d_seg.as<uint32_t>(); tb.ent = d_ent.as<uint2>(); tb.dinv = d_dinv.as<uint32_t>();
            tb.wr = d_wr.as<uint8_t>(); tb.vt = d_vt.as<uint8_t>(); tb.plane = tri_plane; tb.kt = kt; tb.nl = nl;
            auto tbk = f.small ? tri_base_kernel<true> : tri_base_kernel<false>;
            GCU_CHECK(cudaFuncSetAttribute(tbk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRI_BASE_SMEM));
            tbk<<<kt / TRI_BASE, TRI_BASE_THREADS, TRI_BASE_SMEM, stream>>>(tb);
            GCU_CHECK(cudaGetLastError());
            ctr.kernel_launches += 1;
        }
        const CUtensorMap map_wr = make_tmap_u8(d_wr.p, T, (uint64_t)nl * kt, TC_BK, TC_BM);
        const CUtensorMap map_vt = make_tmap_u8(d_vt.p, T, (uint64_t)nl * kt, TC_BK, bn);
        const CUtensorMap map_adt = make_tmap_u8(d_adt.p, T, (uint64_t)nl * kt, TC_BK, bn);
        for (uint32_t s = TRI_BASE; s < T; s *= 2) {
            const uint32_t nb = kt / (2 * s);
            const CUtensorMap map_gx = make_tmap_u8(d_gx.p, s, (uint64_t)nl * (kt / 2), TC_BK, TC_BM);
            TcArgs ga{};          // G = P' Q
            ga.f = f; ga.mp = kt; ga.np = kt; ga.n_rows = s; ga.n_cols = s; ga.m_tiles = s / TC_BM; ga.n_tiles = s / bn;
            ga.kchunk = tkc; ga.k_end = s; ga.k_mod = T; ga.tri = 1;
            ga.a_row0 = 0; ga.a_row_b = 2 * s; ga.a_k0 = 0; ga.a_k_b = 2 * s;
            ga.b_row0 = s; ga.b_row_b = 2 * s; ga.b_k0 = 0; ga.b_k_b = 2 * s;
            ga.o1 = d_gx.as<uint8_t>(); ga.o1_plane = (uint64_t)(kt / 2) * s; ga.o1_ld = s; ga.o1_row_b = s;
            launch_limb_gemm<EPI_LIMBS>(f.small != 0, map_wr, map_adt, ga, nb, stream);
            TcArgs ha{};          // H = -(G R'), written into Wr and (transposed) Vt
            ha.f = f; ha.mp = kt / 2; ha.np = kt; ha.n_rows = s; ha.n_cols = s; ha.m_tiles = s / TC_BM; ha.n_tiles = s / bn;
            ha.kchunk = tkc; ha.k_end = s; ha.k_mod = T; ha.neg = 1; ha.tri = 2;
            ha.a_row0 = 0; ha.a_row_b = s; ha.a_k0 = 0; ha.a_k_b = 0;
            ha.b_row0 = s; ha.b_row_b = 2 * s; ha.b_k0 = s; ha.b_k_b = 2 * s;
            ha.o1 = d_wr.as<uint8_t>(); ha.o1_plane = tri_plane; ha.o1_ld = T; ha.o1_row0 = 0; ha.o1_row_b = 2 * s; ha.o1_col0 = s; ha.o1_col_b = 2 * s;
            ha.o2 = d_vt.as<uint8_t>(); ha.o2_plane = tri_plane; ha.o2_ld = T; ha.o2_row0 = s; ha.o2_row_b = 2 * s; ha.o2_col0 = 0; ha.o2_col_b = 2 * s;
            launch_limb_gemm<EPI_LIMBS>(f.small != 0, map_gx, map_vt, ha, nb, stream);
            ctr.kernel_launches += 2;
        }
        GCU_CHECK(cudaEventRecord(ev[8], stream));
        // 2. tile by tile: cd = C_t + M[:, < tT] A[< tT, tile t], then M[:, tile t] = -cd Inv_t
        const CUtensorMap map_mx = make_tmap_u8(d_mx.p, kp, (uint64_t)nl * mp, TC_BK, TC_BM);
        const CUtensorMap map_cx = make_tmap_u8(d_cx.p, T, (uint64_t)nl * mp, TC_BK, TC_BM);
        const CUtensorMap map_bs = bs ? make_tmap_u8(d_bs_at.p, TC_BK, (uint64_t)nl * (bs_at_plane / 128), TC_BK, bn) : map_cx;
        TriRowsArgs ra{};
        ra.L = L; ra.seg = d_seg.as<uint32_t>(); ra.ent = d_ent.as<uint2>(); ra.row_begin = rank; ra.row_step = world; ra.n_local = n_local;
        const uint32_t* fpiv = mc_k ? nullptr : d_first.as<uint32_t>();      // combinations all start at mc_first
        if (mc_k && mp > n_local) GCU_CHECK(cudaMemsetAsync(d_cd.as<uint32_t>() + (uint64_t)n_local * T, 0, (uint64_t)(mp - n_local) * T * sizeof(uint32_t), stream));
        for (uint32_t t = 0; t < ntp; ++t) {
            const uint32_t tw = lay_tile_width(L, t);
            if (mc_k) {      // C_t of the combinations: a column block of the dense rows
                GCU_CHECK(cudaMemcpy2DAsync(d_cd.p, (size_t)T * sizeof(uint32_t), d_mc_c.as<uint32_t>() + (uint64_t)t * T, mc_ldc * sizeof(uint32_t),
                                            (size_t)T * sizeof(uint32_t), n_local, cudaMemcpyDeviceToDevice, stream));
            } else {
                ra.dst = d_cd.as<uint32_t>(); ra.ld = T; ra.width = T; ra.t_lo = t; ra.t_hi = t + 1; ra.col_base = t * T; ra.n_pad = mp;
                tri_rows_kernel<<<mp, 128, 0, stream>>>(ra);
                GCU_CHECK(cudaGetLastError());
            }
            TcArgs ta{};
            ta.f = f; ta.kp = kp; ta.mp = mp; ta.np = T; ta.n_rows = n_local; ta.n_cols = tw; ta.ld = T;
            ta.d = d_cd.as<uint32_t>(); ta.m_tiles = mp / TC_BM; ta.n_tiles = T / bn;
            ta.first_piv = fpiv; ta.k_first = mc_first; ta.row_begin = rank; ta.row_step = world;
            ta.kchunk = kchunk; ta.k_end = t * T;
            ta.o1 = d_cx.as<uint8_t>(); ta.o1_plane = (uint64_t)mp * T; ta.o1_ld = T;
            if (bs) {
                if (t) {      // the occupied blocks of A[< tT, tile t], block-compressed, into the one-tile buffer
                    oa_at.z0 = t; oa_at.id0 = bs_tile_base[t];
                    opk<<<dim3(t * (T / OP_KB), chunks, 1), 256, op_sm, stream>>>(oa_at);
                    GCU_CHECK(cudaGetLastError());
                    ctr.kernel_launches += 1;
                }
                ta.np = (uint32_t)(bs_at_plane / 128); ta.klist = d_bs_klist.as<uint32_t>(); ta.kl_base = d_bs_base.as<uint32_t>();
                ta.kl_cb0 = t * per_tile; ta.kl_id0 = bs_tile_base[t];
            }
            const CUtensorMap map_at = bs ? map_bs : t ? make_tmap_u8(d_at.as<uint8_t>() + at_off[t], (uint64_t)t * T, (uint64_t)nl * T, TC_BK, bn) : map_vt;
            launch_limb_gemm<EPI_RMW_LIMBS>(f.small != 0, map_mx, map_at, ta, 1, stream);
            TcArgs xa{};
            xa.f = f; xa.mp = mp; xa.np = kt; xa.n_rows = n_local; xa.n_cols = tw; xa.m_tiles = mp / TC_BM; xa.n_tiles = (tw + bn - 1) / bn;
            xa.first_piv = fpiv; xa.k_first = mc_first; xa.row_begin = rank; xa.row_step = world;
            xa.kmin_sub = t * T; xa.skip_hi = (t + 1) * T;
            xa.kchunk = tkc; xa.k_end = (tw + TC_BK - 1) / TC_BK * TC_BK;
            xa.b_row0 = t * T; xa.neg = 1; xa.tri = 2;
            xa.o1 = d_mx.as<uint8_t>(); xa.o1_plane = mx_plane; xa.o1_ld = kp; xa.o1_col0 = t * T;
            xa.stats = ea.stats; xa.cnt_row_ptr = s_row_ptr;
            launch_limb_gemm<EPI_LIMBS>(f.small != 0, map_cx, map_vt, xa, 1, stream);
            ctr.kernel_launches += 3;
        }
        // 3. D scattered into D' (the product below adds M B)
        if (L.n_nonpiv && !mc_k) {
            ra.dst = dloc; ra.ld = L.ld_dprime; ra.width = L.ld_dprime; ra.t_lo = ntp; ra.t_hi = L.n_tiles; ra.col_base = L.n_top; ra.n_pad = n_local;
            tri_rows_kernel<<<n_local, 128, 0, stream>>>(ra);
            GCU_CHECK(cudaGetLastError());
            ctr.kernel_launches += 1;
        }
    } else if (block) {
        // R rows per CTA, 2 CTAs per SM (kernels_elim_block.cuh); D is scattered into the zeroed D' block
        const uint32_t R = f.small ? BlkRows<true>::R : BlkRows<false>::R;
        const size_t bsmem = f.small ? blk_smem_bytes<true>(L.n_windows) : blk_smem_bytes<false>(L.n_windows);
        auto bk = f.small ? elim_block_kernel<true> : elim_block_kernel<false>;
        GCU_CHECK(cudaFuncSetAttribute(bk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
        int bocc = 0;
        GCU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bocc, bk, BLK_THREADS, bsmem));
        if (bocc < 1) throw std::runtime_error("blocked elimination kernel does not fit on an SM");
        const uint32_t n_blocks = (n_local + R - 1) / R;
        const int bgrid = std::max(1, std::min<int>(prop.multiProcessorCount * bocc, (int)n_blocks));
        d_mt.ensure((uint64_t)bgrid * std::max(1u, L.n_windows) * R * 32 * sizeof(uint32_t));
        if (!mc_k) GCU_CHECK(cudaMemsetAsync(dloc, 0, (uint64_t)n_local * L.ld_dprime * sizeof(uint32_t), stream));
        BlkArgs b{};
        b.L = L; b.f = f; b.mu32 = ea.mu32; b.negp = ea.negp; b.norm_every = ea.norm_every;
        b.seg = ea.seg; b.ent = ea.ent; b.diag = ea.diag; b.dnz = ea.dnz; b.first_piv = ea.first_piv;
        b.dprime = dloc; b.mt = d_mt.as<uint32_t>(); b.queue = ea.queue;
        b.row_begin = rank; b.row_step = world; b.n_local = n_local; b.row_ptr = ea.row_ptr; b.stats = ea.stats;
        b.mx = ea.mx; b.mx_plane = ea.mx_plane; b.mx_kp = ea.mx_kp; b.mx_nl = ea.mx_nl;
        if (mc_k) { b.dense_c = d_mc_c.as<uint32_t>(); b.dense_ld = mc_ldc; b.mc_first = mc_first; }
        bk<<<bgrid, BLK_THREADS, bsmem, stream>>>(b);
        GCU_CHECK(cudaGetLastError());
        ctr.kernel_launches += 1;
    } else if (n_local && L.n_nonpiv) {
        kern<<<elim_grid, elim_threads, smem, stream>>>(ea);
        GCU_CHECK(cudaGetLastError());
        ctr.kernel_launches += 1;
    }
    if (!panel) GCU_CHECK(cudaEventRecord(ev[8], stream));
    GCU_CHECK(cudaEventRecord(ev[7], stream));
    if (schur) {
        // D' = D + M B (kernels_schur_tc.cuh): the limb planes as 2-D byte tensors [NL * rows][kp]
        TcArgs ta{};
        ta.f = f; ta.kp = kp; ta.mp = mp; ta.np = np; ta.n_rows = n_local; ta.n_cols = L.n_nonpiv; ta.ld = L.ld_dprime;
        ta.d = dloc; ta.row_nz = nzloc; ta.m_tiles = mp / TC_BM; ta.n_tiles = np / bn;
        ta.first_piv = mc_k ? nullptr : d_first.as<uint32_t>(); ta.k_first = mc_first; ta.row_begin = rank; ta.row_step = world;
        ta.kchunk = kchunk; ta.k_end = k_used;
        const CUtensorMap map_mx = make_tmap_u8(d_mx.p, kp, (uint64_t)nl * mp, TC_BK, TC_BM);
        if (bs) { ta.np = (uint32_t)(bs_plane / 128); ta.klist = d_bs_klist.as<uint32_t>(); ta.kl_base = d_bs_base.as<uint32_t>(); ta.kl_cb0 = n_cb_piv; ta.kl_id0 = (uint32_t)nb_a; }
        const CUtensorMap map_bt = bs ? (bs_plane ? make_tmap_u8(d_bs_op.p, TC_BK, (uint64_t)nl * (bs_plane / 128), TC_BK, bn) : map_mx)
                                      : make_tmap_u8(d_bt.p, kp, (uint64_t)nl * np, TC_BK, bn);
        launch_limb_gemm<EPI_RMW>(f.small != 0, map_mx, map_bt, ta, 1, stream);
        ctr.kernel_launches += 1;
        ctr.schur_macs = (uint64_t)n_local * L.n_top * L.n_nonpiv;
        // every dense product of the stage, in mod-p multiply-adds (x nl^2 = int8 multiply-adds on the tensor cores)
        ctr.dense_macs = bs ? (uint64_t)((double)(nb_total - nb_a) * 128.0 * bn * mp) : ctr.schur_macs;
        if (panel) {
            double m = bs ? (double)nb_a * 128.0 * bn * mp + (double)mp * T * (double)kt : 0.0;
            for (uint32_t t = 0; t < ntp && !bs; ++t) m += (double)mp * T * ((double)t * T + T);
            for (uint32_t s = TRI_BASE; s < T; s *= 2) m += 2.0 * (double)(kt / (2 * s)) * s * s * s;
            ctr.dense_macs += (uint64_t)m;
        }
    } else { ctr.schur_macs = 0; ctr.dense_macs = 0; }
    GCU_CHECK(cudaEventRecord(ev[3], stream));
    eliminated = true;
}

void gamba_cuda_engine::exchange() {
    if (cfg.world_size <= 1) return;
    if (!gather_fn) throw std::runtime_error("world_size > 1 but no all-gather callback set");
    const size_t bytes_d = (size_t)n_local_max * L.ld_dprime * sizeof(uint32_t);
    const size_t bytes_z = (size_t)n_local_max * sizeof(uint32_t);
    char* recv_d = d_dprime.as<char>() + (size_t)mc_base * L.ld_dprime * sizeof(uint32_t);      // behind the combinations of earlier draws
    char* recv_z = d_row_nz.as<char>() + (size_t)mc_base * sizeof(uint32_t);
    const char* send_d = recv_d + (size_t)cfg.rank * bytes_d;
    const char* send_z = recv_z + (size_t)cfg.rank * bytes_z;
    if (bytes_d && gather_fn(gather_user, send_d, recv_d, bytes_d, stream)) throw std::runtime_error("all-gather of D' failed");
    if (bytes_z && gather_fn(gather_user, send_z, recv_z, bytes_z, stream)) throw std::runtime_error("all-gather of row flags failed");
}

void gamba_cuda_engine::echelon(gamba_cuda_step_out* out, uint32_t mc_rows) {
    if (!eliminated) throw std::runtime_error("eliminate() has not run");
    uint32_t nr = n_gather_rows;
    const uint32_t nc = L.n_nonpiv;
    uint32_t* ech_d = d_dprime.as<uint32_t>();
    uint32_t* ech_nz = d_row_nz.as<uint32_t>();
    uint32_t npiv = 0;
    uint64_t nnz_new = 0;
    d_counts.ensure(256);
    // Monte-Carlo stage: buffers before the timed region
    const uint32_t mnl = f.small ? 2u : 4u, mbn = f.small ? TcCfg<2>::BN : TcCfg<4>::BN;
    const uint32_t rp = (nr + 127u) / 128u * 128u, kp_ = (mc_rows + 127u) / 128u * 128u, np_ = (nc + 127u) / 128u * 128u;
    if (mc_rows) {
        try {
            d_mc_s.ensure((uint64_t)mnl * kp_ * rp);
            d_mc_dt.ensure((uint64_t)mnl * np_ * rp);
            d_mc_dc.ensure((uint64_t)kp_ * L.ld_dprime * sizeof(uint32_t));
            d_mc_nz.ensure((uint64_t)kp_ * sizeof(uint32_t));
        } catch (std::exception const&) { mc_rows = 0; mc_fell_back = true; }     // no room for the combinations: the rows themselves
    }
    GCU_CHECK(cudaMemsetAsync(d_counts.p, 0, 32, stream));
    GCU_CHECK(cudaMemsetAsync(d_counts.as<char>() + 32, 0xff, 32, stream));
    GCU_CHECK(cudaEventRecord(ev[4], stream));
    if (mc_rows && nr && nc) {
        // Dc = S D' on the tensor cores; the echelon then sees mc_rows rows
        McArgs ma{};
        ma.n_rows = nr; ma.n_cols = nc; ma.ld = L.ld_dprime; ma.d = d_dprime.as<uint32_t>(); ma.row_nz = d_row_nz.as<uint32_t>();
        ma.dt = d_mc_dt.as<uint8_t>(); ma.s = d_mc_s.as<uint8_t>(); ma.dt_plane = (uint64_t)np_ * rp; ma.s_plane = (uint64_t)kp_ * rp;
        ma.rp = rp; ma.k = mc_rows; ma.nl = mnl; ma.seed = cfg.seed * 0x9e3779b97f4a7c15ull + 0x5bd1e995u * (++mc_ech_draw);
        GCU_CHECK(cudaMemsetAsync(d_mc_s.p, 0, (uint64_t)mnl * kp_ * rp, stream));
        GCU_CHECK(cudaMemsetAsync(d_mc_dt.p, 0, (uint64_t)mnl * np_ * rp, stream));
        GCU_CHECK(cudaMemsetAsync(d_mc_dc.p, 0, (uint64_t)kp_ * L.ld_dprime * sizeof(uint32_t), stream));
        GCU_CHECK(cudaMemsetAsync(d_mc_nz.p, 0, (uint64_t)kp_ * sizeof(uint32_t), stream));
        mc_dt_kernel<<<dim3((nc + 31) / 32, rp / 128), 256, 0, stream>>>(ma);
        GCU_CHECK(cudaGetLastError());
        mc_s_kernel<<<(uint32_t)(((uint64_t)mc_rows * (rp / 4) + 255) / 256), 256, 0, stream>>>(ma);
        GCU_CHECK(cudaGetLastError());
        TcArgs ta{};
        ta.f = f; ta.kp = rp; ta.mp = kp_; ta.np = np_; ta.n_rows = mc_rows; ta.n_cols = nc; ta.ld = L.ld_dprime;
        ta.d = d_mc_dc.as<uint32_t>(); ta.row_nz = d_mc_nz.as<uint32_t>(); ta.m_tiles = kp_ / TC_BM; ta.n_tiles = np_ / mbn;
        ta.first_piv = nullptr; ta.k_first = 0; ta.row_begin = 0; ta.row_step = 1;
        ta.kchunk = f.small ? TcCfg<2>::KCHUNK : TcCfg<4>::KCHUNK; ta.k_end = rp;
        const CUtensorMap map_s = make_tmap_u8(d_mc_s.p, rp, (uint64_t)mnl * kp_, TC_BK, TC_BM);
        const CUtensorMap map_dt = make_tmap_u8(d_mc_dt.p, rp, (uint64_t)mnl * np_, TC_BK, mbn);
        launch_limb_gemm<EPI_RMW>(f.small != 0, map_s, map_dt, ta, 1, stream);
        ctr.kernel_launches += 3;
        nr = mc_rows; ech_d = d_mc_dc.as<uint32_t>(); ech_nz = d_mc_nz.as<uint32_t>();
    }
    if (nr && nc) {
        d_lead.ensure((size_t)3 * nr * sizeof(uint32_t));
        d_nz_list.ensure((size_t)nr * sizeof(uint32_t));
        d_piv_row.ensure((size_t)nr * sizeof(uint32_t));
        d_piv_col.ensure((size_t)nr * sizeof(uint32_t));
        d_is_piv.ensure(nc);
        d_out_ptr.ensure((size_t)(nr + 1) * sizeof(uint64_t));
        EchArgs ea{};
        ea.f = f; ea.n_rows = nr; ea.n_cols = nc; ea.ld = L.ld_dprime;
        ea.d = ech_d; ea.row_nz = ech_nz;
        ea.lead = d_lead.as<uint32_t>(); ea.inv_lead = d_lead.as<uint32_t>() + (size_t)2 * nr; ea.nz_list = d_nz_list.as<uint32_t>();
        ea.piv_row = d_piv_row.as<uint32_t>(); ea.piv_col = d_piv_col.as<uint32_t>();
        ea.is_piv_col = d_is_piv.as<uint8_t>(); ea.counts = d_counts.as<uint32_t>();
        ea.out_ptr = d_out_ptr.as<uint64_t>();
        {
            // blocked Gauss-Jordan, trailing updates on the tensor cores (kernels_echelon_mma.cuh)
            const uint32_t nl = f.small ? 2u : 4u;
            const uint32_t ncp = (nc + EM_CW - 1) / EM_CW * EM_CW;
            d_lead.ensure((size_t)2 * nr * sizeof(uint32_t));
            d_piv_row.ensure((size_t)2 * nr * sizeof(uint32_t));
            d_piv_col.ensure((size_t)2 * nr * sizeof(uint32_t));
            d_em_live.ensure((size_t)nr * sizeof(uint32_t));
            d_em_xa.ensure((size_t)nr * nl * 32);
            d_em_pbt.ensure((size_t)nl * ncp * 32);
            EmArgs em{};
            em.f = f; em.n_rows = nr; em.n_cols = nc; em.ld = L.ld_dprime; em.ncp = ncp;
            em.d = ea.d; em.row_nz = ea.row_nz; em.lead = d_lead.as<uint32_t>(); em.live = d_em_live.as<uint32_t>();
            em.xa = d_em_xa.as<uint8_t>(); em.pbt = d_em_pbt.as<uint8_t>();
            em.piv_row = d_piv_row.as<uint32_t>(); em.piv_col = d_piv_col.as<uint32_t>();
            em.is_piv_col = ea.is_piv_col; em.counts = ea.counts; em.out_ptr = ea.out_ptr;
            auto kern = f.small ? (void*)echelon_mma_kernel<2> : (void*)echelon_mma_kernel<4>;
            GCU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EM_SMEM));
            int occ = 0;
            GCU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, EM_THREADS, EM_SMEM));
            if (occ < 1) throw std::runtime_error("tensor-core echelon kernel does not fit on an SM");
            void* args[] = {&em};
            GCU_CHECK(cudaLaunchCooperativeKernel(kern, dim3(prop.multiProcessorCount), dim3(EM_THREADS), args, EM_SMEM, stream));
        }
        ctr.kernel_launches += 1;
        h_misc.ensure(64);
        GCU_CHECK(cudaMemcpyAsync(h_misc.p, d_counts.p, 8, cudaMemcpyDeviceToHost, stream));
        GCU_CHECK(cudaStreamSynchronize(stream));
        npiv = h_misc.as<uint32_t>()[1];
        if (npiv) {
            h_out_ptr.ensure((size_t)(npiv + 1) * sizeof(uint64_t));
            GCU_CHECK(cudaMemcpyAsync(h_out_ptr.p, d_out_ptr.p, (size_t)(npiv + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
            GCU_CHECK(cudaStreamSynchronize(stream));
            nnz_new = h_out_ptr.as<uint64_t>()[npiv];
            d_out_col.ensure(nnz_new * sizeof(uint32_t));
            d_out_val.ensure(nnz_new * sizeof(uint32_t));
            EmitArgs em{};
            em.f = f; em.n_cols = nc; em.n_piv = npiv; em.ld = L.ld_dprime; em.d = ech_d;
            em.piv_row = d_piv_row.as<uint32_t>(); em.piv_col = d_piv_col.as<uint32_t>();
            em.is_piv_col = d_is_piv.as<uint8_t>(); em.out_ptr = d_out_ptr.as<uint64_t>();
            em.out_col = d_out_col.as<uint32_t>(); em.out_val = d_out_val.as<uint32_t>();
            emit_kernel<<<npiv, 256, 0, stream>>>(em);
            GCU_CHECK(cudaGetLastError());
            ctr.kernel_launches += 1;
        }
    }
    GCU_CHECK(cudaEventRecord(ev[5], stream));

    r_row_ptr.assign(npiv + 1, 0);
    r_col.resize(nnz_new); r_val32.resize(nnz_new); r_piv.resize(npiv);
    if (npiv) {
        std::memcpy(r_row_ptr.data(), h_out_ptr.p, (size_t)(npiv + 1) * sizeof(uint64_t));
        GCU_CHECK(cudaMemcpyAsync(r_col.data(), d_out_col.p, nnz_new * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        GCU_CHECK(cudaMemcpyAsync(r_val32.data(), d_out_val.p, nnz_new * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        GCU_CHECK(cudaMemcpyAsync(r_piv.data(), d_piv_col.p, (size_t)npiv * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    }
    h_misc.ensure(64);
    GCU_CHECK(cudaMemcpyAsync(h_misc.as<char>() + 16, d_queue_stats.as<char>() + 16, 24, cudaMemcpyDeviceToHost, stream));
    GCU_CHECK(cudaStreamSynchronize(stream));
    float ms = 0;
    GCU_CHECK(cudaEventElapsedTime(&ms, ev[2], ev[3])); t_elim_dev = ms * 1e-3;
    GCU_CHECK(cudaEventElapsedTime(&ms, ev[6], ev[7])); t_sparse_dev = ms * 1e-3;
    GCU_CHECK(cudaEventElapsedTime(&ms, ev[6], ev[8])); t_inv_dev = (use_schur && use_panel && n_local) ? ms * 1e-3 : 0.0;
    ctr.seconds_inverse_last = t_inv_dev;
    GCU_CHECK(cudaEventElapsedTime(&ms, ev[2], ev[6])); ctr.seconds_operands_last = ms * 1e-3;
    ctr.solve_mode = (use_schur && n_local) ? (use_panel ? 2u : use_block ? 1u : 0u) : 0u;
    ctr.panel_t = ctr.solve_mode == 2 ? PANEL_T : 0u;
    t_schur_dev = t_elim_dev - t_sparse_dev;
    ctr.seconds_sparse_last = t_sparse_dev; ctr.seconds_schur_last = t_schur_dev;
    GCU_CHECK(cudaEventElapsedTime(&ms, ev[4], ev[5])); t_ech_dev = ms * 1e-3;
    ctr.pivots_applied = reinterpret_cast<unsigned long long*>(h_misc.as<char>() + 16)[0];
    ctr.fma_applied = reinterpret_cast<unsigned long long*>(h_misc.as<char>() + 16)[1];
    if (ctr.solve_mode == 2) ctr.fma_applied = ctr.dense_macs - ctr.schur_macs;      // the solve is dense: what the products did
    ctr.fma_top = reinterpret_cast<unsigned long long*>(h_misc.as<char>() + 16)[2];
    if (n_local && L.n_top && !mc_k) {
        last_density = (double)ctr.pivots_applied / ((double)n_local * L.n_top); last_a = cur_a; have_history = true;
        if (use_schur && !use_block && !use_panel && last_a > 0 && ctr.pivots_applied > 1000)      // the sparse kernel ran: its real work
            hitw = std::min(16.0, std::max(0.25, (double)ctr.fma_applied / ((double)ctr.pivots_applied * last_a * 0.5)));
        if (use_schur && L.n_top >= 32768 && n_local >= 64 && t_sparse_dev > 5e-3) {      // calibrate the model of the mode that ran (steps big enough
                                                                                           // for the model's terms, not its fixed costs, to matter)
            const uint32_t nl_ = f.small ? 2u : 4u;
            const double mpad = (double)((n_local + 127u) / 128u * 128u);
            const int m = use_panel ? 2 : use_block ? 1 : 0;
            const double model = m == 2 ? model_panel(mpad, L.n_top, nl_, PANEL_T, opt_bsparse && L.n_top >= 8192 ? last_occ : 1.0)
                               : m == 1 ? model_block(n_local, L.n_top, cur_a, f.small != 0)
                                        : hitw * model_sparse(n_local, last_density, L.n_top, cur_a, f.small != 0);
            if (model > 0) calib[m] = std::min(3.0, std::max(0.33, 0.5 * calib[m] + 0.5 * t_sparse_dev / model));
        }
    }
    ctr.steps += 1;

    if ((flags & GAMBA_CUDA_STEP_FINAL) && npiv != L.n_bot)
        throw std::runtime_error("final inter-reduction: rank of D' differs from the number of rows");

    out->n_new = npiv; out->n_nonpiv = nc; out->nnz_new = nnz_new;
    out->row_ptr = r_row_ptr.data(); out->col_rel = r_col.data(); out->piv_rel = r_piv.data();
    switch (cfg.coeff_bytes) {
        case 4: out->coeff = r_val32.data(); break;
        case 2: r_val16.resize(nnz_new); for (uint64_t i = 0; i < nnz_new; ++i) r_val16[i] = (uint16_t)r_val32[i]; out->coeff = r_val16.data(); break;
        default: r_val8.resize(nnz_new); for (uint64_t i = 0; i < nnz_new; ++i) r_val8[i] = (uint8_t)r_val32[i]; out->coeff = r_val8.data(); break;
    }
    out->zero_reductions = L.n_bot - std::min(L.n_bot, npiv);
    out->seconds_eliminate = t_elim_dev; out->seconds_echelon = t_ech_dev;
    out->seconds_device = t_prep_dev + t_elim_dev + t_ech_dev;
    out->h2d_bytes = h2d_bytes;
    out->d2h_bytes = (npiv + 1) * 8ull + nnz_new * 8ull + npiv * 4ull + 24;
}

// Exact mode: every row of C|D is eliminated. Monte-Carlo mode (kernels_mc.cuh, chosen in choose_layout): K random combinations
// of all the rows are eliminated instead and the result is accepted only if the rank of D' leaves a margin of 32 combinations
// (K - rank >= 32): K uniform vectors of a space of dimension rank span it except with probability ~p^-(K - rank) <= 3^-32, so
// the row space - hence the canonical RREF - is that of the exact computation. When the margin is not met the reduced
// combinations stay where they are in D' (they are rows of the row space whatever the echelon did to them) and as many fresh
// ones are added behind them (the total doubles); when half the rows would be reached, the step runs exactly.
void gamba_cuda_engine::run_reduce(gamba_cuda_step_out* out) {
    uint32_t k = mc_plan;
    const uint32_t world = std::max(1u, cfg.world_size);
    mc_base = 0;
    mc_cap_rows = k ? L.n_bot / 2 + 32 * world + 8 : 0;    // the draws stop at n_bot / 2 combinations; each is padded to a multiple of the ranks
    double acc_elim = 0, acc_ech = 0;
    ctr.mc_attempts = 0;
    uint32_t combos = 0;
    for (;;) {
        mc_k = k; ++mc_draw; ++ctr.mc_attempts;
        double tq = wall_s();
        eliminate();
        exchange();
        dbg_[5] += wall_s() - tq; tq = wall_s();
        // Monte-Carlo combinations in front of the echelon: K = predicted rank + margin combinations of the rows of D'
        // instead of the rows (most of which reduce to zero); accepted only if K - rank >= 32, else the rows themselves
        uint32_t kk = 0;
        if (opt_mc_ech && !mc_k && !(flags & GAMBA_CUDA_STEP_FINAL) && last_rank >= 0 && n_gather_rows >= 512 && L.n_nonpiv) {
            const uint32_t want = ((uint32_t)(last_rank + 32 + last_rank / 8) + 127u) / 128u * 128u;
            if ((uint64_t)want * 2 <= n_gather_rows) kk = want;
        }
        mc_fell_back = false;
        echelon(out, kk);
        if (mc_fell_back) kk = 0;
        if (kk && out->n_new + 32 > kk) {      // margin not met
            const double t_first = t_ech_dev;
            kk = 0;
            echelon(out, 0);
            t_ech_dev += t_first; ctr.steps -= 1;
        }
        mc_ech_k = kk; ctr.mc_ech_combinations = kk;
        dbg_[6] += wall_s() - tq;
        acc_elim += t_elim_dev; acc_ech += t_ech_dev;
        if (mc_k) {
            combos += mc_k;
            if (out->n_new + 32 <= combos) break;                    // accepted
            ctr.steps -= 1;
            const uint64_t next = combos;                            // as many again
            if ((combos + next) * 2 > L.n_bot) { k = 0; mc_base = 0; combos = 0; }                   // exact
            else { k = (uint32_t)next; mc_base = n_gather_rows; }
            continue;
        }
        break;
    }
    ctr.mc_combinations = combos;
    mc_base = 0;
    last_rank = out->n_new;
    out->seconds_eliminate = acc_elim; out->seconds_echelon = acc_ech;
    out->seconds_device = t_prep_dev + acc_elim + acc_ech;
}

// ---- row generation of symbolic preprocessing on the device (kernels_rowgen.cuh) -------------------------------------

// grow a device buffer keeping its first `used` bytes
static void grow_keep(gcu::DevBuf& b, size_t used, size_t want, cudaStream_t st) {
    if (want <= b.cap) return;
    if (b.va) { b.ensure(want); return; }          // virtual-memory mode: grows in place
    gcu::DevBuf bigger;
    bigger.ensure(std::max(want, b.cap * 2));
    if (used) GCU_CHECK(cudaMemcpyAsync(bigger.p, b.p, used, cudaMemcpyDeviceToDevice, st));
    std::swap(bigger.p, b.p); std::swap(bigger.cap, b.cap);      // the old block goes back to the pool, ordered on the stream
    std::swap(bigger.va, b.va); std::swap(bigger.va_size, b.va_size); std::swap(bigger.chunks, b.chunks);
}

void gamba_cuda_engine::rg_begin(uint32_t stride) {
    if (stride == 0 || ((stride + 1u) & ~1u) > (uint32_t)RG_MAX_SD) throw std::runtime_error("rowgen: unsupported exponent-vector stride");
    if (stride != rg.stride) { rg.stride = stride; rg.sd = (stride + 1u) & ~1u; rg.terms = 0; rg.n_basis = 0; }
    rg.count = 0; rg.n_rows_entries = 0;
    rg.cnt.ensure(16);
    GCU_CHECK(cudaMemsetAsync(rg.cnt.p, 0, 16, stream));
    if (rg.n_slots) GCU_CHECK(cudaMemsetAsync(rg.slots.p, 0, (size_t)rg.n_slots * sizeof(uint32_t), stream));
}

// append `count` monomials of the driver's basis table (ex: [count][stride], hv: hashes); *first = index of the first one
void gamba_cuda_engine::rg_add_monomials(uint32_t count, const uint16_t* ex, const uint32_t* hv, uint64_t* first) {
    if (!rg.sd) throw std::runtime_error("rowgen: begin has not been called");
    const uint32_t sd = rg.sd, st = rg.stride;
    *first = rg.n_basis;
    if (!count) return;
    grow_keep(rg.bex, rg.n_basis * sd * sizeof(uint16_t), (rg.n_basis + count) * sd * sizeof(uint16_t), stream);
    grow_keep(rg.bhv, rg.n_basis * sizeof(uint32_t), (rg.n_basis + count) * sizeof(uint32_t), stream);
    if (sd == st) {
        GCU_CHECK(cudaMemcpyAsync(rg.bex.as<uint16_t>() + rg.n_basis * sd, ex, (size_t)count * sd * sizeof(uint16_t), cudaMemcpyHostToDevice, stream));
    } else {
        const size_t bytes = (size_t)count * sd * sizeof(uint16_t);
        GCU_CHECK(cudaStreamSynchronize(stream));            // the staging buffer may still be the source of the previous upload
        rg.stage.ensure(bytes);
        uint16_t* se = rg.stage.as<uint16_t>();
        for (uint32_t k = 0; k < count; ++k) { std::memcpy(se + (size_t)k * sd, ex + (size_t)k * st, st * sizeof(uint16_t)); se[(size_t)k * sd + st] = 0; }
        GCU_CHECK(cudaMemcpyAsync(rg.bex.as<uint16_t>() + rg.n_basis * sd, se, bytes, cudaMemcpyHostToDevice, stream));
    }
    GCU_CHECK(cudaMemcpyAsync(rg.bhv.as<uint32_t>() + rg.n_basis, hv, (size_t)count * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    GCU_CHECK(cudaStreamSynchronize(stream));                // the caller's arrays may move (its table grows)
    rg.n_basis += count;
}

// n_polys polynomials in one upload: poly i = n_terms[i] basis-monomial handles at mon[i]; first[i] = where it starts in the term store
void gamba_cuda_engine::rg_add_polys(uint32_t n_polys, const uint32_t* n_terms, const uint32_t* const* mon, uint64_t* first) {
    if (!rg.sd) throw std::runtime_error("rowgen: begin has not been called");
    uint64_t n = 0;
    for (uint32_t i = 0; i < n_polys; ++i) { first[i] = rg.terms + n; n += n_terms[i]; }
    if (!n) return;
    grow_keep(rg.pmon, rg.terms * sizeof(uint32_t), (rg.terms + n) * sizeof(uint32_t), stream);
    GCU_CHECK(cudaStreamSynchronize(stream));            // the staging buffer may still be the source of the previous upload
    rg.stage.ensure(n * sizeof(uint32_t));
    uint32_t* sm = rg.stage.as<uint32_t>();
    uint64_t o = 0;
    for (uint32_t i = 0; i < n_polys; ++i) {
        for (uint32_t k = 0; k < n_terms[i]; ++k) if (mon[i][k] >= rg.n_basis) throw std::runtime_error("rowgen: polynomial names a monomial that was not uploaded");
        std::memcpy(sm + o, mon[i], (size_t)n_terms[i] * sizeof(uint32_t));
        o += n_terms[i];
    }
    GCU_CHECK(cudaMemcpyAsync(rg.pmon.as<uint32_t>() + rg.terms, sm, n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    rg.terms += n;
}

void gamba_cuda_engine::rg_add_rows(uint32_t n_rows, const uint64_t* first_term, const uint32_t* n_terms, const uint16_t* q_ex,
                                    const uint32_t* q_hash, uint32_t* lead_out, uint32_t* n_monomials) {
    if (!rg.sd) throw std::runtime_error("rowgen: begin has not been called");
    const uint32_t sd = rg.sd, st = rg.stride;
    // Sub-batches: the table must have room for every term of a sub-batch being new (it cannot grow inside the kernel), while
    // almost all terms are hits - so a sub-batch is capped near the table's size (at least 4 M terms).
    for (uint32_t r0 = 0; r0 < n_rows;) {
        const uint64_t cap_terms = std::max<uint64_t>((uint64_t)1 << 22, rg.count);
        uint32_t r1 = r0; uint64_t total = 0;
        while (r1 < n_rows && (r1 == r0 || total + n_terms[r1] <= cap_terms)) total += n_terms[r1++];
        const uint32_t nb = r1 - r0;
        if ((uint64_t)rg.count + total >= 0xfffffff0ull) throw std::runtime_error("rowgen: too many monomials for 32-bit handles");
        const uint32_t need = rg.count + (uint32_t)total;
        if (need > rg.cap) {
            const uint32_t cap = std::max<uint32_t>(need, rg.cap + rg.cap / 2 + (1u << 16));
            grow_keep(rg.hv, (size_t)rg.count * sizeof(uint32_t), (size_t)cap * sizeof(uint32_t), stream);
            grow_keep(rg.ex, (size_t)rg.count * sd * sizeof(uint16_t), (size_t)cap * sd * sizeof(uint16_t), stream);
            rg.cap = cap;
        }
        if ((uint64_t)need * 2 > rg.n_slots) {
            uint64_t ns = std::max<uint64_t>(rg.n_slots, (uint64_t)1 << 20);
            while ((uint64_t)need * 2 > ns) ns *= 2;
            if (ns > 0x80000000ull) throw std::runtime_error("rowgen: monomial table too large");
            rg.slots.ensure(ns * sizeof(uint32_t));
            rg.n_slots = (uint32_t)ns;
            GCU_CHECK(cudaMemsetAsync(rg.slots.p, 0, ns * sizeof(uint32_t), stream));
            if (rg.count) {
                rowgen_rehash_kernel<<<(rg.count + 255) / 256, 256, 0, stream>>>(rg.slots.as<uint32_t>(), rg.n_slots - 1, rg.hv.as<uint32_t>(), rg.count);
                GCU_CHECK(cudaGetLastError());
                ctr.kernel_launches += 1;
            }
        }
        grow_keep(rg.rows, rg.n_rows_entries * sizeof(uint32_t), (rg.n_rows_entries + total) * sizeof(uint32_t), stream);
        // batch arrays through one page-locked staging buffer
        const size_t o_first = 0, o_off = o_first + (size_t)nb * 8, o_q = o_off + (size_t)(nb + 1) * 8, o_hq = o_q + (size_t)nb * sd * 2,
                     o_end = (o_hq + (size_t)nb * 4 + 15) & ~(size_t)15;
        GCU_CHECK(cudaStreamSynchronize(stream));
        rg.stage.ensure(o_end);
        char* sb = rg.stage.as<char>();
        std::memcpy(sb + o_first, first_term + r0, (size_t)nb * 8);
        uint64_t* off = reinterpret_cast<uint64_t*>(sb + o_off);
        off[0] = rg.n_rows_entries;
        for (uint32_t i = 0; i < nb; ++i) off[i + 1] = off[i] + n_terms[r0 + i];
        uint16_t* sq = reinterpret_cast<uint16_t*>(sb + o_q);
        for (uint32_t i = 0; i < nb; ++i) {
            std::memcpy(sq + (size_t)i * sd, q_ex + (size_t)(r0 + i) * st, st * sizeof(uint16_t));
            if (sd != st) sq[(size_t)i * sd + st] = 0;
        }
        std::memcpy(sb + o_hq, q_hash + r0, (size_t)nb * 4);
        rg.b_first.ensure(o_end);
        GCU_CHECK(cudaMemcpyAsync(rg.b_first.p, sb, o_end, cudaMemcpyHostToDevice, stream));
        rg.b_lead.ensure((size_t)nb * 4);
        RowGenArgs a{};
        a.sd = sd; a.bex = rg.bex.as<uint16_t>(); a.bhv = rg.bhv.as<uint32_t>(); a.pmon = rg.pmon.as<uint32_t>(); a.n_rows = nb;
        a.first_term = reinterpret_cast<const uint64_t*>(rg.b_first.as<char>() + o_first);
        a.row_off = reinterpret_cast<const uint64_t*>(rg.b_first.as<char>() + o_off);
        a.q = reinterpret_cast<const uint16_t*>(rg.b_first.as<char>() + o_q);
        a.hq = reinterpret_cast<const uint32_t*>(rg.b_first.as<char>() + o_hq);
        a.slots = rg.slots.as<uint32_t>(); a.mask = rg.n_slots - 1; a.hv = rg.hv.as<uint32_t>(); a.ex = rg.ex.as<uint16_t>();
        a.count = rg.cnt.as<uint32_t>(); a.cap = rg.cap; a.overflow = rg.cnt.as<uint32_t>() + 1;
        a.row_mon = rg.rows.as<uint32_t>(); a.lead = rg.b_lead.as<uint32_t>();
        const uint32_t grid = std::min<uint32_t>(nb, (uint32_t)prop.multiProcessorCount * 8);
        if (sd <= 16) rowgen_kernel<8><<<grid, RG_THREADS, 0, stream>>>(a);
        else if (sd <= 32) rowgen_kernel<16><<<grid, RG_THREADS, 0, stream>>>(a);
        else rowgen_kernel<64><<<grid, RG_THREADS, 0, stream>>>(a);
        GCU_CHECK(cudaGetLastError());
        ctr.kernel_launches += 1;
        uint32_t res[2] = {0, 0};
        GCU_CHECK(cudaMemcpyAsync(res, rg.cnt.p, 8, cudaMemcpyDeviceToHost, stream));
        GCU_CHECK(cudaMemcpyAsync(lead_out + r0, rg.b_lead.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, stream));
        GCU_CHECK(cudaStreamSynchronize(stream));
        if (res[1]) throw std::runtime_error("rowgen: monomial table overflow");
        rg.count = res[0];
        rg.n_rows_entries += total;
        r0 = r1;
    }
    *n_monomials = rg.count;
}

void gamba_cuda_engine::rg_fetch(uint32_t first, uint32_t count, uint16_t* ex_out, uint32_t* hv_out) {
    if ((uint64_t)first + count > rg.count) throw std::runtime_error("rowgen: monomial range out of bounds");
    if (!count) return;
    const uint32_t sd = rg.sd, st = rg.stride;
    const size_t bytes_ex = (size_t)count * sd * sizeof(uint16_t);
    GCU_CHECK(cudaStreamSynchronize(stream));
    rg.stage.ensure(bytes_ex);
    GCU_CHECK(cudaMemcpyAsync(rg.stage.p, rg.ex.as<uint16_t>() + (size_t)first * sd, bytes_ex, cudaMemcpyDeviceToHost, stream));
    GCU_CHECK(cudaMemcpyAsync(hv_out, rg.hv.as<uint32_t>() + first, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    GCU_CHECK(cudaStreamSynchronize(stream));
    const uint16_t* se = rg.stage.as<uint16_t>();
    for (uint32_t k = 0; k < count; ++k) std::memcpy(ex_out + (size_t)k * st, se + (size_t)k * sd, st * sizeof(uint16_t));
}

// ------------------------------------------------------------------ C ABI -----

template <class F>
static int guarded(F&& fn) {
    try { fn(); return 0; }
    catch (std::exception const& e) { gcu::g_last_error = e.what(); return 1; }
    catch (...) { gcu::g_last_error = "unknown error"; return 2; }
}

extern "C" {

int gamba_cuda_abi_version(void) { return GAMBA_CUDA_ABI_VERSION; }
const char* gamba_cuda_last_error(void) { return gcu::g_last_error.c_str(); }

int gamba_cuda_create(const gamba_cuda_config* cfg, gamba_cuda_engine** out) {
    return guarded([&] {
        if (!cfg || !out) throw std::runtime_error("null argument");
        if (cfg->p < 3 || cfg->p > 0x7fffffffu || (cfg->p & 1) == 0) throw std::runtime_error("p must be an odd prime < 2^31");
        if (cfg->coeff_bytes != 1 && cfg->coeff_bytes != 2 && cfg->coeff_bytes != 4) throw std::runtime_error("coeff_bytes must be 1, 2 or 4");
        if (cfg->coeff_bytes < 4 && cfg->p >= (1ull << (8 * cfg->coeff_bytes))) throw std::runtime_error("p does not fit coeff_bytes");
        int ndev = 0;
        GCU_CHECK(cudaGetDeviceCount(&ndev));
        if (ndev <= 0 || (int)cfg->device >= ndev) throw std::runtime_error("no such CUDA device");
        GCU_CHECK(cudaSetDevice((int)cfg->device));
        auto* e = new gamba_cuda_engine();
        try {
            e->cfg = *cfg;
            if (e->cfg.world_size == 0) e->cfg.world_size = 1;
            const uint32_t mont = cfg->repr == GAMBA_CUDA_REPR_MONTGOMERY ? (cfg->coeff_bytes == 4 ? 64u : 32u) : 0u;
            e->f = gcu::fp_make(cfg->p, mont);
            GCU_CHECK(cudaGetDeviceProperties(&e->prop, (int)cfg->device));
            if (e->prop.major < 10) throw std::runtime_error("this library is built for sm_100a (B200) only");
            GCU_CHECK(cudaDeviceGetAttribute(&e->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, (int)cfg->device));
            GCU_CHECK(cudaDeviceGetAttribute(&e->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, (int)cfg->device));
            GCU_CHECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = (int)cfg->device;
            GCU_CHECK(cudaMemPoolCreate(&e->pool, &props));
            uint64_t keep = ~0ull;                       // freed blocks stay in the pool for the next step
            GCU_CHECK(cudaMemPoolSetAttribute(e->pool, cudaMemPoolAttrReleaseThreshold, &keep));
            // Put some memory into the engine's pool up front (one allocation, released straight back to the pool): the first steps
            // of a run then grow their buffers without a trip to the driver each time (1 s of a 3 s cyclic9-15 run otherwise).
            // GAMBA_POOL_PREWARM_GB (default 1 - the big buffers are virtual-memory ranges, not pool blocks; capped at a quarter of
            // the free memory; 0 = off). The pool dies with the engine.
            {
                size_t free_b = 0, total_b = 0;
                GCU_CHECK(cudaMemGetInfo(&free_b, &total_b));
                const char* pw = std::getenv("GAMBA_POOL_PREWARM_GB");
                const double want_gb = pw ? std::atof(pw) : 1.0;
                const size_t want = (size_t)std::min<double>(want_gb * 1e9, (double)free_b / 4);
                if (want > ((size_t)64 << 20)) {
                    void* warm = nullptr;
                    if (cudaMallocFromPoolAsync(&warm, want, e->pool, e->stream) == cudaSuccess) GCU_CHECK(cudaFreeAsync(warm, e->stream));
                    else cudaGetLastError();
                    GCU_CHECK(cudaStreamSynchronize(e->stream));
                }
            }
            for (auto& ev : e->ev) GCU_CHECK(cudaEventCreate(&ev));
            e->ctr.sm_count = (uint32_t)e->prop.multiProcessorCount;
        } catch (...) { gcu::g_alloc_stream = e->stream; gcu::g_alloc_pool = e->pool; const cudaStream_t st = e->stream; const cudaMemPool_t pl = e->pool; delete e; gamba_cuda_engine::release_device(st, pl); throw; }
        *out = e;
    });
}

void gamba_cuda_destroy(gamba_cuda_engine* e) {
    if (!e) return;
    cudaSetDevice((int)e->cfg.device);
    gcu::g_alloc_stream = e->stream; gcu::g_alloc_pool = e->pool;       // the buffers go back to the engine's pool, then the pool goes
    const cudaStream_t st = e->stream; const cudaMemPool_t pl = e->pool;
    if (st) cudaStreamSynchronize(st);                                      // mapped ranges are unmapped by the buffers' destructors
    delete e;
    gamba_cuda_engine::release_device(st, pl);
}

void* gamba_cuda_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void gamba_cuda_host_free(void* p) { if (p) cudaFreeHost(p); }

int gamba_cuda_stage(gamba_cuda_engine* e, const gamba_cuda_step_in* in) {
    return guarded([&] { GCU_CHECK(cudaSetDevice((int)e->cfg.device)); gcu::g_alloc_stream = e->stream; gcu::g_alloc_pool = e->pool; e->stage(in); });
}

int gamba_cuda_reduce_staged(gamba_cuda_engine* e, gamba_cuda_step_out* out) {
    return guarded([&] {
        GCU_CHECK(cudaSetDevice((int)e->cfg.device));
        gcu::g_alloc_stream = e->stream; gcu::g_alloc_pool = e->pool;
        const double t0 = gcu::wall_s();
        e->run_reduce(out);
        out->seconds_total = gcu::wall_s() - t0;
    });
}

int gamba_cuda_reduce(gamba_cuda_engine* e, const gamba_cuda_step_in* in, gamba_cuda_step_out* out) {
    return guarded([&] {
        GCU_CHECK(cudaSetDevice((int)e->cfg.device));
        gcu::g_alloc_stream = e->stream; gcu::g_alloc_pool = e->pool;
        const double t0 = gcu::wall_s();
        e->stage(in);
        e->run_reduce(out);
        out->seconds_total = gcu::wall_s() - t0;
    });
}

int gamba_cuda_set_allgather(gamba_cuda_engine* e, gamba_cuda_allgather_fn fn, void* user) {
    return guarded([&] { e->gather_fn = fn; e->gather_user = user; });
}

int gamba_cuda_set_broadcast(gamba_cuda_engine* e, gamba_cuda_broadcast_fn fn, void* user) {
    return guarded([&] { e->bcast_fn = fn; e->bcast_user = user; });
}

#define GCU_ENTER(e) GCU_CHECK(cudaSetDevice((int)(e)->cfg.device)); gcu::g_alloc_stream = (e)->stream; gcu::g_alloc_pool = (e)->pool

int gamba_cuda_rowgen_begin(gamba_cuda_engine* e, uint32_t stride) {
    return guarded([&] { GCU_ENTER(e); e->rg_begin(stride); });
}
int gamba_cuda_rowgen_add_monomials(gamba_cuda_engine* e, uint32_t count, const uint16_t* ex, const uint32_t* hash, uint64_t* first) {
    return guarded([&] { GCU_ENTER(e); e->rg_add_monomials(count, ex, hash, first); });
}
int gamba_cuda_rowgen_add_polys(gamba_cuda_engine* e, uint32_t n_polys, const uint32_t* n_terms, const uint32_t* const* mon, uint64_t* first_term) {
    return guarded([&] { GCU_ENTER(e); e->rg_add_polys(n_polys, n_terms, mon, first_term); });
}
int gamba_cuda_rowgen_add_rows(gamba_cuda_engine* e, uint32_t n_rows, const uint64_t* first_term, const uint32_t* n_terms,
                               const uint16_t* q_ex, const uint32_t* q_hash, uint32_t* lead_out, uint32_t* n_monomials) {
    return guarded([&] { GCU_ENTER(e); e->rg_add_rows(n_rows, first_term, n_terms, q_ex, q_hash, lead_out, n_monomials); });
}
int gamba_cuda_rowgen_fetch_monomials(gamba_cuda_engine* e, uint32_t first, uint32_t count, uint16_t* ex_out, uint32_t* hash_out) {
    return guarded([&] { GCU_ENTER(e); e->rg_fetch(first, count, ex_out, hash_out); });
}
int gamba_cuda_rowgen_rows(gamba_cuda_engine* e, const uint32_t** dev_rows, uint64_t* n_entries) {
    return guarded([&] { *dev_rows = e->rg.rows.as<uint32_t>(); *n_entries = e->rg.n_rows_entries; });
}
int gamba_cuda_rowgen_download_rows(gamba_cuda_engine* e, uint64_t first, uint64_t count, uint32_t* out) {
    return guarded([&] {
        GCU_ENTER(e);
        if (first + count > e->rg.n_rows_entries) throw std::runtime_error("rowgen: row range out of bounds");
        if (count) GCU_CHECK(cudaMemcpyAsync(out, e->rg.rows.as<uint32_t>() + first, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
        GCU_CHECK(cudaStreamSynchronize(e->stream));
    });
}

int gamba_cuda_measure_int8_peak(uint32_t device, uint32_t n, double* tmacs_per_s) {
    return guarded([&] {
        if (!tmacs_per_s || (n != 64 && n != 128 && n != 256)) throw std::runtime_error("n must be 64, 128 or 256");
        GCU_CHECK(cudaSetDevice((int)device));
        cudaDeviceProp prop{};
        GCU_CHECK(cudaGetDeviceProperties(&prop, (int)device));
        if (prop.major < 10) throw std::runtime_error("this library is built for sm_100a (B200) only");
        cudaStream_t st; cudaEvent_t e0, e1;
        GCU_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        GCU_CHECK(cudaEventCreate(&e0)); GCU_CHECK(cudaEventCreate(&e1));
        const uint32_t iters = 20000, grid = (uint32_t)prop.multiProcessorCount;
        auto kern = n == 64 ? gcu::tc_peak_kernel<64> : n == 128 ? gcu::tc_peak_kernel<128> : gcu::tc_peak_kernel<256>;
        const size_t smem = gcu::tc_peak_smem(n);
        GCU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {          // first repetition = warm-up
            GCU_CHECK(cudaEventRecord(e0, st));
            kern<<<grid, 128, smem, st>>>(iters);
            GCU_CHECK(cudaGetLastError());
            GCU_CHECK(cudaEventRecord(e1, st));
            GCU_CHECK(cudaStreamSynchronize(st));
            float ms = 0; GCU_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) best = std::min(best, ms);
        }
        *tmacs_per_s = (double)grid * iters * 4.0 * 128.0 * n * 32.0 / (best * 1e-3) / 1e12;
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
    });
}

int gamba_cuda_get_counters(gamba_cuda_engine* e, gamba_cuda_counters* c) {
    return guarded([&] { *c = e->ctr; });
}

int gamba_cuda_set_option(gamba_cuda_engine* e, const char* key, int64_t value) {
    return guarded([&] {
        const std::string k = key ? key : "";
        if (k == "tile_cols") e->opt_tile_cols = value;
        else if (k == "ctas_per_sm") e->opt_ctas_per_sm = value;
        else if (k == "elim_warps") e->opt_elim_warps = value;
        else if (k == "coeff_cache") { e->opt_coeff_cache = value; e->cache.clear(); e->cache_used = 0; }
        else if (k == "schur") e->opt_schur = value;
        else if (k == "block") e->opt_block = value;
        else if (k == "panel") e->opt_panel = value;
        else if (k == "panel_t") { if (value && (value < 128 || value > 8192 || (value & (value - 1)))) throw std::runtime_error("panel_t must be 128 * 2^j, at most 8192"); e->opt_panel_t = value; }
        else if (k == "panel_max_top") e->opt_panel_max_top = value;
        else if (k == "panel_max_gb") e->opt_panel_max_gb = value;
        else if (k == "bsparse") e->opt_bsparse = value;
        else if (k == "schur_max_gb") e->opt_schur_max_gb = value;
        else if (k == "schur_min_top") e->opt_schur_min_top = value;
        else if (k == "schur_kchunk") e->opt_schur_kchunk = value;
        else if (k == "mc_ech") e->opt_mc_ech = value;
        else if (k == "mc") e->opt_mc = value;
        else if (k == "mc_k") e->opt_mc_k = value;
        else if (k == "mc_min_top") e->opt_mc_min_top = value;
        else if (k == "mc_min_rows") e->opt_mc_min_rows = value;
        else if (k == "rank_hint") e->last_rank = value;
        else throw std::runtime_error("unknown option " + k);
    });
}

}  // extern "C"
