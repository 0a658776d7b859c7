/*
 * gamba_cuda.h — C ABI of the B200 linear-algebra engine for GamBa's F4.
 *
 * This is the drop-in boundary for the reference's two closed engine classes:
 *
 *   linalg_v3<Coeff,Order>(field, seed)     ctor      source/f4.hpp:66
 *   linalg_v3::initialize(matrix_f4&)                 source/f4.hpp:91
 *   linalg_v3::reduce(params) -> seconds              source/f4.hpp:93
 *     (output members read by matrix_f4::extract_new_rows, source/matrix.hpp:799-895)
 *   linalg<Matrix>(matrix) / ::reduce(matrix,params)  source/reduce.hpp:51-52
 *
 * The reference calls C++ templates, not a C ABI; the thin adapter that maps those
 * class interfaces onto the functions below is shown in INTEGRATION.md. The
 * north star fixes the name gamba_cuda_reduce; the other names are ours.
 *
 * One step matrix is handed over in Faugere-Lachartre block form:
 *   rows 0 .. n_top-1            pivot rows A|B; row i is monic with its lead at column i
 *   rows n_top .. n_top+n_bot-1  rows to reduce C|D, in the driver's order
 *   columns 0 .. n_top-1         pivot columns, in ascending original (= decreasing monomial) order
 *   columns n_top .. n_cols-1    non-pivot columns, original relative order kept
 * A row's col_idx list is in the polynomial's term order (lead first), so entry k
 * pairs with element k of the row's SHARED coefficient array coeff_ptr[coeff_id[row]];
 * within a row the A-part (< n_top) and the B-part (>= n_top) are each strictly increasing.
 *
 * Number representation (create flag): GAMBA_CUDA_REPR_STANDARD = residues in [0,p);
 * GAMBA_CUDA_REPR_MONTGOMERY = the reference's field_traits form (source/field.hpp:41-51):
 * x*R mod p with R = 2^32 for 1- and 2-byte coefficients, R = 2^64 for 4-byte ones.
 * Output uses the same representation and element width as the input.
 *
 * Result: the canonical reduced row echelon form of D' = D - C A^-1 B over the
 * non-pivot columns: rows monic, fully inter-reduced, sorted by ascending pivot
 * (= decreasing leading monomial, the order source/basis.hpp:725-727 relies on).
 *
 * Threading: one handle per host thread; calls on a handle are strictly sequential
 * (the reference is single-threaded). All functions return 0 on
 * success and a non-zero code otherwise; gamba_cuda_last_error() gives the message.
 * There is no CPU fallback: without a usable CUDA device gamba_cuda_create fails.
 */
#ifndef GAMBA_CUDA_H
#define GAMBA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAMBA_CUDA_ABI_VERSION 4   /* 4: gamba_cuda_counters grew (dense_macs, seconds_inverse_last, solve_mode, panel_t, seconds_operands_last),
                                      GAMBA_CUDA_STEP_DEVICE_INDICES / _COLIDX, gamba_cuda_measure_int8_peak, gamba_cuda_rowgen_* */

#define GAMBA_CUDA_REPR_STANDARD   0u
#define GAMBA_CUDA_REPR_MONTGOMERY 1u

/* gamba_cuda_step_in.flags */
#define GAMBA_CUDA_STEP_FINAL      1u  /* final inter-reduction (linalg<M>::reduce, reduce.hpp:51-52): the rows
                                          to reduce have distinct non-pivot leads, so the engine checks
                                          n_new == n_bot; same output form (ascending pivot = the bottom
                                          rows' order reversed, matrix.hpp:733-739)                    */

#define GAMBA_CUDA_STEP_DEVICE_INDICES 2u /* row_ptr, col_idx, coeff_id (and row_src, col_map) point to memory of the engine's
                                          device and are used where they are until the next step is staged (a caller that
                                          builds or keeps its step matrices on the GPU; bench.py's resident-input leg).
                                          coeff_ptr / coeff_len stay host arrays (see "coeff_cache"). Not with a broadcast callback. */

#define GAMBA_CUDA_STEP_DEVICE_COLIDX 4u  /* col_idx alone points to device memory: the handle array gamba_cuda_rowgen_add_rows
                                          filled (gamba_cuda_rowgen_rows); every other array is a host array */

typedef struct gamba_cuda_engine gamba_cuda_engine;

typedef struct gamba_cuda_config {
    uint32_t p;             /* field characteristic, odd prime < 2^31                          */
    uint32_t coeff_bytes;   /* element width of coefficient arrays in and out: 1, 2 or 4       */
    uint32_t repr;          /* GAMBA_CUDA_REPR_*                                                */
    uint32_t device;        /* CUDA device ordinal                                              */
    uint64_t seed;          /* seed of the Monte-Carlo combinations (params.seed, params.hpp:47) */
    /* row-sharding of C|D across ranks (one process per GPU). With world_size > 1 the caller
       provides the exchange callbacks below (the Python host wires them to torch.distributed /
       NCCL); with world_size == 1 they are ignored. */
    uint32_t rank, world_size;
} gamba_cuda_config;

typedef struct gamba_cuda_step_in {
    uint32_t n_top, n_bot, n_cols, n_coeff_arrays;
    uint64_t nnz;
    const uint64_t* row_ptr;           /* [n_top + n_bot + 1]                                 */
    const uint32_t* col_idx;           /* [nnz] block-form columns                            */
    const uint32_t* coeff_id;          /* [n_top + n_bot] -> coefficient array of the row     */
    const void* const* coeff_ptr;      /* [n_coeff_arrays], elements coeff_bytes wide         */
    const uint32_t* coeff_len;         /* [n_coeff_arrays]                                    */
    uint32_t flags;
    /* Optional indirect form (all zero / NULL = the direct form above). It lets a driver hand its rows over
       where symbolic preprocessing left them (matrix_f4 keeps rows as monomial handles and numbers the
       columns afterwards, source/matrix.hpp:603-673) without rewriting O(nnz) data on the host:
         row_src  [n_top+n_bot]  row r's entries are col_idx[row_src[r] .. row_src[r] + len(r)), with
                                 len(r) = row_ptr[r+1] - row_ptr[r]; rows may lie anywhere, in any order
         col_map  [col_map_len]  col_idx holds caller handles h; the block-form column is col_map[h]
         col_idx_len             number of entries of col_idx to copy (>= nnz when rows are not packed)  */
    uint32_t col_map_len;
    const uint64_t* row_src;
    const uint32_t* col_map;
    uint64_t col_idx_len;
} gamba_cuda_step_in;

typedef struct gamba_cuda_step_out {
    uint32_t n_new;                    /* rank of D' (== n_bot for GAMBA_CUDA_STEP_FINAL)     */
    uint32_t n_nonpiv;                 /* n_cols - n_top                                      */
    uint64_t nnz_new;
    /* engine-owned host buffers, valid until the next call on the handle */
    const uint64_t* row_ptr;           /* [n_new + 1]                                         */
    const uint32_t* col_rel;           /* [nnz_new] relative non-pivot columns, pivot first   */
    const void* coeff;                 /* [nnz_new] coeff_bytes wide, pivot entry == one      */
    const uint32_t* piv_rel;           /* [n_new] ascending                                   */
    uint64_t zero_reductions;          /* rows of C|D that reduced to zero (stats.zero_reductions) */
    double seconds_total;              /* wall time of the call: H2D + kernels + D2H (f4.hpp:93)   */
    double seconds_device;             /* CUDA-event time of the kernels only                      */
    double seconds_eliminate;          /* ... of the C|D-by-A|B elimination kernel                 */
    double seconds_echelon;            /* ... of the D' echelon                                    */
    uint64_t h2d_bytes, d2h_bytes;
} gamba_cuda_step_out;

int gamba_cuda_abi_version(void);
const char* gamba_cuda_last_error(void);

int gamba_cuda_create(const gamba_cuda_config* cfg, gamba_cuda_engine** out);
void gamba_cuda_destroy(gamba_cuda_engine* e);

/* Page-locked host memory for the caller's step arrays (col_idx above all): H2D copies from it are plain
   DMA transfers. Optional - any host memory is accepted by the calls below. NULL on failure. */
void* gamba_cuda_host_alloc(size_t bytes);
void gamba_cuda_host_free(void* p);

/* One F4 step (or the final inter-reduction): host buffers in, host buffers out. */
int gamba_cuda_reduce(gamba_cuda_engine* e, const gamba_cuda_step_in* in, gamba_cuda_step_out* out);

/* The same split in two, so a caller (and the benchmark) can keep a step resident in HBM:
   gamba_cuda_stage copies the step to the device; gamba_cuda_reduce_staged runs the kernels on
   the staged step (repeatable) and copies the result back. */
int gamba_cuda_stage(gamba_cuda_engine* e, const gamba_cuda_step_in* in);
int gamba_cuda_reduce_staged(gamba_cuda_engine* e, gamba_cuda_step_out* out);

/* Row-block sharding. With world_size > 1 every rank stages the same step,
   eliminates only the rows r with r % world_size == rank, and the reduced rows are exchanged
   through the caller-provided all-gather (NCCL over NVLink in the Python host; gloo on CPU tests).
   gather(user, sendbuf_dev, recvbuf_dev, bytes_per_rank, stream) must behave like ncclAllGather
   on device pointers enqueued on `stream` (a cudaStream_t). */
typedef int (*gamba_cuda_allgather_fn)(void* user, const void* send_dev, void* recv_dev,
                                       size_t bytes_per_rank, void* stream);
int gamba_cuda_set_allgather(gamba_cuda_engine* e, gamba_cuda_allgather_fn fn, void* user);

/* Replication of the step ("ncclBroadcast of the flattened A|B"). With world_size > 1 and
   this callback set, only rank 0 copies the step from its host buffers; gamba_cuda_stage / gamba_cuda_reduce
   on the other ranks take in == NULL and receive the flattened arrays (and their sizes) device to device.
   bcast(user, buf_dev, bytes, root, stream) must behave like ncclBroadcast in place on `stream`. Without the
   callback every rank passes the same step and uploads it itself. */
typedef int (*gamba_cuda_broadcast_fn)(void* user, void* buf_dev, size_t bytes, int root, void* stream);
int gamba_cuda_set_broadcast(gamba_cuda_engine* e, gamba_cuda_broadcast_fn fn, void* user);

/* Introspection used by the benchmark and the tests. */
typedef struct gamba_cuda_counters {
    uint64_t kernel_launches;          /* kernels of this library launched since create         */
    uint64_t steps;
    uint32_t sm_count;
    uint32_t tile_cols;                /* accumulator tile width used by the last step          */
    uint32_t n_tiles;
    uint32_t acc_bits;                 /* 32 or 64                                              */
    uint64_t pivots_applied;           /* sum over rows of pivot rows applied in the last step  */
    uint64_t fma_applied;              /* multiply-adds the elimination kernel did, last step   */
    uint64_t fma_top;                  /* matrix invariant of SURVEY.md §8(d): sum over pivot hits of
                                          nnz(pivot row), this rank's rows, last step              */
    uint32_t mc_combinations;          /* Monte-Carlo combinations eliminated by the last step instead of its rows
                                          (0 = the step ran on the rows themselves)                 */
    uint32_t mc_attempts;              /* ... and how many draws it took to meet the rank margin (the exact run counts) */
    double seconds_stage_last;         /* wall time of the last staging: H2D (or broadcast) + staging kernels */
    double seconds_sparse_last;        /* CUDA-event time of elim_kernel alone (the sparse triangular solve), last step */
    double seconds_schur_last;         /* ... of the tensor-core product D' = D + M B (limb planes + schur kernel)     */
    uint64_t schur_macs;               /* mod-p multiply-adds of that product: rows x n_top x n_nonpiv (0 = sparse path) */
    uint32_t mc_ech_combinations;      /* random combinations of the rows of D' the last echelon ran on (0 = the rows themselves) */
    uint32_t solve_mode;               /* how the multipliers M = -C A^-1 of the last step were found: 0 sparse row-per-CTA kernel,
                                          1 sparse kernel for blocks of rows, 2 blocked triangular solve on the tensor cores       */
    uint64_t dense_macs;               /* mod-p multiply-adds of ALL tensor-core products of the elimination stage, last step
                                          (x 4 / x 16 int8 multiply-adds for 2 / 4 limbs)                                         */
    double seconds_inverse_last;       /* CUDA-event time of the inverse diagonal tiles (inside seconds_sparse_last), last step   */
    uint32_t panel_t;                  /* pivot tile of the blocked solve (0 = not used)                                          */
    uint32_t reserved0;
    double seconds_operands_last;      /* CUDA-event time of building the limb-plane operands of the products (inside
                                          seconds_schur_last), last step                                                          */
} gamba_cuda_counters;
int gamba_cuda_get_counters(gamba_cuda_engine* e, gamba_cuda_counters* c);

/* Row generation of symbolic preprocessing on the device (role of matrix_f4::create_multiplied_poly_matrix_row
   and the matrix' monomial hash table, source/matrix.hpp:387-444; callers insert_spairs :215-347 and symbolic_preprocessing :469-566).
   The step's monomial table and its rows (as 32-bit monomial handles, numbered in order of first appearance) live on the device;
   the driver keeps the reducer search (source/matrix.hpp:569-600) and gets back, per batch, the new monomials and one lead handle
   per row. Exponent vectors are `stride` uint16 wide (whatever layout the driver uses; only equality and addition matter) with
   additive 32-bit hashes: hash(a * b) = hash(a) + hash(b) (source/monomial.hpp:80-90,126-131).
     rowgen_begin          new step: empty table, no rows (basis monomials and polynomials uploaded earlier stay)
     rowgen_add_monomials  append monomials of the driver's basis table (the one its polynomials are stored in: handles, as
                           source/basis.hpp keeps them); *first = index the first of them got
     rowgen_add_polys      upload basis polynomials once, as lists of indices into those monomials; first_term[i] identifies
                           polynomial i in add_rows
     rowgen_add_rows       append the rows poly[i] * x^q[i]; lead_out[i] = handle of row i's first term; *n_monomials = size of the
                           step's table afterwards (handles [old size, new size) are the monomials this batch introduced)
     rowgen_fetch_monomials   exponent vectors and hashes of the step-table handles [first, first + count)
     rowgen_rows           device pointer and length of the handle array: pass it as col_idx with GAMBA_CUDA_STEP_DEVICE_COLIDX,
                           row_src[r] = where row r starts in it (rows are stored in the order they were added)
     rowgen_download_rows  copy a range of the handle array to the host (dumps, debugging)                                    */
int gamba_cuda_rowgen_begin(gamba_cuda_engine* e, uint32_t stride);
int gamba_cuda_rowgen_add_monomials(gamba_cuda_engine* e, uint32_t count, const uint16_t* ex, const uint32_t* hash, uint64_t* first);
int gamba_cuda_rowgen_add_polys(gamba_cuda_engine* e, uint32_t n_polys, const uint32_t* n_terms, const uint32_t* const* mon, uint64_t* first_term);
int gamba_cuda_rowgen_add_rows(gamba_cuda_engine* e, uint32_t n_rows, const uint64_t* first_term, const uint32_t* n_terms,
                               const uint16_t* q_ex, const uint32_t* q_hash, uint32_t* lead_out, uint32_t* n_monomials);
int gamba_cuda_rowgen_fetch_monomials(gamba_cuda_engine* e, uint32_t first, uint32_t count, uint16_t* ex_out, uint32_t* hash_out);
int gamba_cuda_rowgen_rows(gamba_cuda_engine* e, const uint32_t** dev_rows, uint64_t* n_entries);
int gamba_cuda_rowgen_download_rows(gamba_cuda_engine* e, uint64_t first, uint64_t count, uint32_t* out);

/* Microbenchmark: the int8 tensor-core rate of `device` with operands resident in shared memory (tcgen05.mma kind::i8,
   128 x n x 32 per instruction, n = 64, 128 or 256, one CTA per SM), in 10^12 multiply-adds per second. bench.py uses it as
   the measured denominator of its tensor roofline. */
int gamba_cuda_measure_int8_peak(uint32_t device, uint32_t n, double* tmacs_per_s);

/* Options: "tile_cols", "ctas_per_sm" (0 = automatic),
   "coeff_cache" (default 0; 1: coefficient arrays stay on the device across steps, keyed by host pointer - the
   caller promises they are immutable while the engine lives, as the reference's basis storage is,
   source/matrix.hpp:189-193; setting the option again drops the cache),
   "schur" (default 1: the B half of the elimination, D' = D + M B, runs as a dense limb-split product on the
   tensor cores whenever its operands fit "schur_max_gb" (default 96) GB of HBM; 0: sparse path for both halves),
   "panel" (with "schur": the multipliers M = -C A^-1 by a blocked triangular solve on the tensor cores, kernels_tri.cuh;
   0 never, 1 (default) when the cost estimate from the previous step says it beats the sparse kernels, 2 always),
   "panel_t" (pivot tile of that solve: 128 * 2^j, 0 = automatic), "panel_max_top",
   "block" (with "schur": the sparse half of the elimination for R rows per CTA from one pass over the pivot rows,
   kernels_elim_block.cuh; 0 never, 1 (default) for steps with >= 32768 pivot rows once the previous step had
   more than half of its multipliers non-zero, 2 always), "elim_warps", "schur_min_top", "schur_kchunk" (tuning / tests),
   "mc_ech" (default 1: the echelon runs on K = predicted rank + margin random combinations of the rows of D', formed on the
   tensor cores, and is accepted only if K - rank >= 32 - otherwise it is repeated on the rows themselves; the result is the
   canonical RREF either way; "rank_hint" sets the prediction),
   "mc" (Monte-Carlo combinations of the rows to reduce, the reference's "Monte-Carlo F4" (README.md:5; the seed of
   linalg_v3{field, seed}, source/f4.hpp:66): K = predicted rank + margin uniform combinations S (C|D) are eliminated
   instead of the n_bot rows, accepted only if K - rank >= 32, topped up with as many again otherwise, and given up for
   the rows themselves at n_bot / 2 - the result is the canonical RREF either way, DESIGN.md §5.7. 0 never; 1 (default)
   on steps with >= "mc_min_top" (default 262144) pivot rows and >= "mc_min_rows" (default 256) rows to reduce whose
   predicted K is at most n_bot / 2; 2 on every step with K <= n_bot / 2 that a
   blocked solve runs on), "mc_k" (first draw, 0 = from the prediction), "rank_hint" (expected rank of D'; default: the
   previous step's).
   fma_top counts the combinations, not the rows, on steps where mc_combinations > 0. */
int gamba_cuda_set_option(gamba_cuda_engine* e, const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* GAMBA_CUDA_H */
