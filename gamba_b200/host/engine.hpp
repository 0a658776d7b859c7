// The linear-algebra boundary as the host driver sees it.
//
// Mirrors the consumer side of the reference's closed engines:
//   linalg_v3<Coeff,Order>::initialize/reduce  (call sites source/f4.hpp:91-93,
//     output contract read by source/matrix.hpp:799-895)
//   linalg<Matrix>::reduce                       (call site source/reduce.hpp:51-52)
// A step matrix is handed over in Faugere-Lachartre block form:
//   * rows 0..n_top-1 are the pivot rows A|B, row i has its (unit) lead at column i;
//   * rows n_top.. are the rows to reduce, C|D, in the driver's order;
//   * columns 0..n_top-1 are the pivot columns (ascending original order),
//     columns n_top..n_cols-1 the non-pivot columns in their original order;
//   * a row's column list is in the polynomial's term order, so entry k pairs
//     with coefficient k of the row's *shared* coefficient array coeff_id[row].
// Coefficients are standard residues in [0,p); every row is monic.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace gb {

// Allocation of the one O(nnz) array that crosses the boundary every step. The CUDA engine installs
// page-locked allocation (gamba_cuda_host_alloc) so that the H2D copy is a plain DMA; default: malloc.
struct HostAllocHook { void* (*alloc)(size_t) = nullptr; void (*release)(void*) = nullptr; };
inline HostAllocHook& host_alloc_hook() { static HostAllocHook h; return h; }

class ColArray {
public:
    ColArray() = default;
    ColArray(const ColArray&) = delete;
    ColArray& operator=(const ColArray&) = delete;
    ~ColArray() { drop(); }
    uint32_t* data() { return p_; }
    const uint32_t* data() const { return p_; }
    size_t size() const { return n_; }
    uint32_t& operator[](size_t i) { return p_[i]; }
    const uint32_t& operator[](size_t i) const { return p_[i]; }
    void resize(size_t n) {          // contents are NOT preserved
        if (n > cap_) {
            drop();
            const size_t cap = n + n / 2 + 1024;
            auto& h = host_alloc_hook();
            p_ = static_cast<uint32_t*>(h.alloc ? h.alloc(cap * sizeof(uint32_t)) : std::malloc(cap * sizeof(uint32_t)));
            if (!p_) throw std::runtime_error("out of host memory for the step matrix");
            pinned_ = h.alloc != nullptr; cap_ = cap;
        }
        n_ = n;
    }
    void assign(const uint32_t* a, const uint32_t* b) { resize((size_t)(b - a)); std::copy(a, b, p_); }

private:
    void drop() {
        if (p_) { if (pinned_) host_alloc_hook().release(p_); else std::free(p_); }
        p_ = nullptr; cap_ = n_ = 0;
    }
    uint32_t* p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
    bool pinned_ = false;
};

struct StepMatrix {
    uint32_t p = 0;
    uint32_t n_top = 0, n_bot = 0, n_cols = 0;
    std::vector<uint64_t> row_ptr;      // n_top + n_bot + 1
    ColArray col;                       // nnz, block-form column numbers
    std::vector<uint32_t> coeff_id;     // per row, index into cf_ptr / cf_len
    std::vector<const uint32_t*> cf_ptr;
    std::vector<uint32_t> cf_len;
    // Indirect form (what the driver produces; gamba_cuda.h): the entries stay where symbolic preprocessing
    // wrote them, as step-table handles; row r's entries are src[row_src[r] ..), column = col_map[handle].
    const uint32_t* src = nullptr;
    uint64_t src_len = 0;
    std::vector<uint64_t> row_src;
    std::vector<uint32_t> col_map;
    bool src_on_device = false;         // src points to device memory (rows generated on the GPU, Engine::rowgen()): not readable here
    uint32_t col_at(uint32_t r, uint64_t j) const { return src ? col_map[src[row_src[r] + j]] : col[row_ptr[r] + j]; }
    uint64_t nnz() const { return row_ptr.empty() ? 0 : row_ptr.back(); }
    uint32_t n_rows() const { return n_top + n_bot; }
    uint32_t n_nonpiv() const { return n_cols - n_top; }
};

// Reduced row echelon form of D' = D - C A^-1 B over the non-pivot columns:
// rows sorted by ascending pivot, monic, fully inter-reduced; col holds
// *relative* non-pivot column numbers (0..n_nonpiv-1), strictly increasing in
// a row, pivot entry first.
struct NewRows {
    uint32_t n_new = 0;
    std::vector<uint64_t> row_ptr;  // n_new + 1
    std::vector<uint32_t> col;
    std::vector<uint32_t> val;
    double seconds = 0.0;           // engine time for this step (f4.hpp:93)
    double kernel_seconds = 0.0;    // device-only part when known
    uint64_t fma_top = 0;           // SURVEY.md §8(d) invariant
    std::string note;               // engine's one-line account of the step (round table, -v 2)
};

// Row generation of symbolic preprocessing on the engine's device (include/gamba_cuda.h: gamba_cuda_rowgen_*; role of
// matrix_f4::create_multiplied_poly_matrix_row and the matrix' monomial table, source/matrix.hpp:387-444): the step's monomial
// table and rows live on the GPU; the driver keeps a mirror of the monomials (exponents, hashes) for its reducer search.
struct RowGenDevice {
    virtual ~RowGenDevice() = default;
    virtual void begin(uint32_t stride) = 0;
    // append monomials of the driver's basis table; returns the index of the first one
    virtual uint64_t add_monomials(uint32_t count, const uint16_t* ex, const uint32_t* hv) = 0;
    // upload polynomials (lists of basis-table handles) in one go; first[i] names poly i in add_rows
    virtual void add_polys(uint32_t n_polys, const uint32_t* n_terms, const uint32_t* const* mon, uint64_t* first) = 0;
    // rows poly[i] * x^q[i]; lead_out[i] = handle of the first term of row i; returns the size of the table afterwards
    virtual uint32_t add_rows(uint32_t n_rows, const uint64_t* first_term, const uint32_t* n_terms, const uint16_t* q_ex,
                              const uint32_t* q_hash, uint32_t* lead_out) = 0;
    virtual void fetch(uint32_t first, uint32_t count, uint16_t* ex_out, uint32_t* hv_out) = 0;
    virtual const uint32_t* rows_device(uint64_t& n_entries) = 0;
    virtual void download_rows(uint64_t first, uint64_t count, uint32_t* out) = 0;
};

struct Engine {
    virtual ~Engine() = default;
    // non-null: the engine generates the step's rows on its device (the CUDA engine); null: the driver does (CPU oracle)
    virtual RowGenDevice* rowgen() { return nullptr; }
    virtual const char* name() const = 0;
    // One F4 step (role of linalg_v3::initialize + reduce).
    virtual void reduce_step(const StepMatrix& m, NewRows& out) = 0;
    // Final inter-reduction (role of linalg<M>::reduce): same computation; the
    // bottom rows have distinct non-pivot leads so n_new == n_bot.
    virtual void reduce_final(const StepMatrix& m, NewRows& out) { reduce_step(m, out); }
};

// ---- step-matrix files (tests, bench, kernel development without the driver) --

struct StepFile {
    StepMatrix m;
    std::vector<uint32_t> pool;       // backing store for m.cf_ptr
    std::vector<uint64_t> pool_off;
};

inline void write_step_file(const std::string& path, const StepMatrix& m) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    uint64_t pool_len = 0;
    for (auto l : m.cf_len) pool_len += l;
    uint64_t hdr[8] = {0x3150545342474aull /* "JGBSTP1" */, m.p, m.n_top, m.n_bot, m.n_cols,
                       m.nnz(), (uint64_t)m.cf_len.size(), pool_len};
    std::fwrite(hdr, 8, 8, f);
    std::fwrite(m.row_ptr.data(), 8, m.row_ptr.size(), f);
    if (m.src) {
        std::vector<uint32_t> tmp;
        for (uint32_t r = 0; r < m.n_rows(); ++r) {
            const uint64_t len = m.row_ptr[r + 1] - m.row_ptr[r];
            tmp.resize(len);
            for (uint64_t j = 0; j < len; ++j) tmp[j] = m.col_at(r, j);
            std::fwrite(tmp.data(), 4, len, f);
        }
    } else {
        std::fwrite(m.col.data(), 4, m.col.size(), f);
    }
    std::fwrite(m.coeff_id.data(), 4, m.coeff_id.size(), f);
    std::fwrite(m.cf_len.data(), 4, m.cf_len.size(), f);
    for (size_t i = 0; i < m.cf_len.size(); ++i) std::fwrite(m.cf_ptr[i], 4, m.cf_len[i], f);
    std::fclose(f);
}

inline void write_rows_file(const std::string& path, const NewRows& r) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    uint64_t hdr[3] = {0x3157524e42474aull /* "JGBNRW1" */, r.n_new, r.row_ptr.empty() ? 0 : r.row_ptr.back()};
    std::fwrite(hdr, 8, 3, f);
    std::fwrite(r.row_ptr.data(), 8, r.row_ptr.size(), f);
    std::fwrite(r.col.data(), 4, r.col.size(), f);
    std::fwrite(r.val.data(), 4, r.val.size(), f);
    std::fclose(f);
}

}  // namespace gb
