/*
 * ug4b200.h — C ABI of the B200-native algebra kernels for ugcore's assembled-matrix
 * linear-solve path (GMG V-cycle preconditioning CG / BiCGStab on CRS matrices).
 *
 * This is the drop-in boundary: the functions a ugcore "GPU" algebra
 * (GPUAlgebra / GPUBlockAlgebra<3>, registered next to CPUAlgebra in
 * ugbase/bridge/util_algebra_dependent.h:63-100) binds instead of the legacy
 *   extern "C" bool CUDA_VecAdd2 / CUDA_VecAdd3 / CUDA_JacobiApply(...)
 *       ugbase/lib_algebra/gpu_algebra/cuda/common_cuda.h:38-51
 *   cusparseDcsrmv / cublasDdot / cublasDnrm2 call sites
 *       ugbase/lib_algebra/gpu_algebra/gpusparsematrix_impl.h:157-164
 *       ugbase/lib_algebra/gpu_algebra/gpuvector.h:200-214, 297-303
 *   CUDAManager (device selection, handles, H2D/D2H helpers)
 *       ugbase/lib_algebra/gpu_algebra/cuda/cuda_manager.h:79-180
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns 0 on success or a
 *     UG4B200_ERR_* code and never throws; ug4b200_last_error() gives the text.
 *     (Legacy wrappers returned `bool true` and relied on CUDA_CHECK_STATUS ->
 *     UG_THROW in the caller, cuda_manager.h:61-77; the C++ wrapper layer turns a
 *     non-zero code into UG_THROW the same way.)
 *   - vectors are raw DEVICE pointers to fp64, `n` counts doubles (block vectors:
 *     n = blocks * block size, block entries contiguous like Vector<DenseVector<..>>).
 *   - all work is stream-ordered on the context's stream; nothing blocks the host
 *     unless the function returns a value to a HOST pointer.
 *   - arithmetic is IEEE fp64 without FMA contraction, summed in ugcore's order
 *     (ascending column index inside a row), so SpMV / Jacobi / AXPY results are
 *     bit-identical to the CPU algebra; only reductions differ (tree vs sequential).
 *   - no CPU fallback: if no CUDA device is usable, ctx_create fails.
 */
#ifndef UG4B200_H
#define UG4B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define UG4B200_VERSION 100

enum {
	UG4B200_OK = 0,
	UG4B200_ERR_CUDA = 1,      /* a CUDA runtime call failed */
	UG4B200_ERR_ARG = 2,       /* invalid argument */
	UG4B200_ERR_NOMEM = 3,
	UG4B200_ERR_NCCL = 4,
	UG4B200_ERR_STATE = 5      /* call not valid in the current state */
};

typedef struct ug4b200_ctx ug4b200_ctx;
typedef struct ug4b200_matrix ug4b200_matrix;
typedef struct ug4b200_interface ug4b200_interface;

/* ------------------------------------------------------------------ context */

/* device: CUDA ordinal (one process/rank per device; replaces CUDAManager::init's
 * "most SMs" pick, cuda_manager.cpp:54-68).  stream: a cudaStream_t created by the
 * caller (e.g. torch's current stream) or NULL to let the context own one. */
int ug4b200_ctx_create(int device, void* stream, ug4b200_ctx** out);
int ug4b200_ctx_destroy(ug4b200_ctx* ctx);
const char* ug4b200_last_error(const ug4b200_ctx* ctx); /* ctx may be NULL: last global error */
int ug4b200_sync(ug4b200_ctx* ctx);
void* ug4b200_stream(ug4b200_ctx* ctx);
/* number of kernels this context launched so far */
int ug4b200_launch_count(const ug4b200_ctx* ctx, int64_t* n);
/* while set, every subsequent kernel of this context returns immediately when
 * *flag != 0 (device int).  NULL clears.  Used to make queued-ahead Krylov
 * iterations no-ops once the device-side convergence check has fired. */
int ug4b200_set_guard(ug4b200_ctx* ctx, const int* dev_flag);

/* Batched small operations.  Calls of the SpMV family, Jacobi steps, element-wise vector
 * operations, gathers and the dense LU solve whose operand has at most `max_rows` rows are not
 * launched one by one: they are recorded in call order and executed by ONE kernel (a single
 * thread-block cluster, cluster barrier between operations) as soon as any other work is
 * enqueued on the context's stream or the host synchronises — stream order is unchanged, only
 * the 4-6 us launch + dependency latency per operation of the multigrid's coarse levels
 * (mg_solver_impl.hpp:1685-1964 on levels of a few thousand rows) collapses into one launch.
 * Results are bit-identical to the stand-alone kernels.  on = 0 turns recording off;
 * max_rows < 0 keeps the current limit (default 16384, UG4B200_BATCH_MAX_ROWS).
 * ug4b200_stream() flushes, so foreign work enqueued on the stream stays ordered. */
int ug4b200_batch_enable(ug4b200_ctx* ctx, int on, int64_t max_rows);
int ug4b200_batch_flush(ug4b200_ctx* ctx);
/* operations executed inside batch kernels so far; cluster size in use (0: unavailable) */
int ug4b200_batch_stats(const ug4b200_ctx* ctx, int64_t* batched_ops, int* cluster_size);

/* CUDA-graph capture of a stream-ordered call sequence (e.g. one Krylov iteration whose
 * scalars all live on the device): begin, issue calls, end -> replay with launch. */
typedef struct ug4b200_graph ug4b200_graph;
int ug4b200_graph_begin(ug4b200_ctx* ctx);
int ug4b200_graph_end(ug4b200_ctx* ctx, ug4b200_graph** out);
int ug4b200_graph_launch(ug4b200_ctx* ctx, ug4b200_graph* g);
int ug4b200_graph_destroy(ug4b200_ctx* ctx, ug4b200_graph* g);

/* events on the context's stream (device-side timing, lagged convergence polling) */
int ug4b200_event_create(ug4b200_ctx* ctx, void** ev);
int ug4b200_event_record(ug4b200_ctx* ctx, void* ev);
int ug4b200_event_sync(ug4b200_ctx* ctx, void* ev);
int ug4b200_event_elapsed_ms(ug4b200_ctx* ctx, void* ev_start, void* ev_stop, float* ms);
int ug4b200_event_destroy(ug4b200_ctx* ctx, void* ev);

/* ------------------------------------------------------------------- memory */

int ug4b200_alloc(ug4b200_ctx* ctx, size_t bytes, void** dptr);
int ug4b200_free(ug4b200_ctx* ctx, void* dptr);
int ug4b200_h2d(ug4b200_ctx* ctx, void* dst, const void* src, size_t bytes);   /* async if src is pinned */
int ug4b200_d2h(ug4b200_ctx* ctx, void* dst, const void* src, size_t bytes);   /* returns after completion */
int ug4b200_d2h_async(ug4b200_ctx* ctx, void* dst_pinned, const void* src, size_t bytes);
int ug4b200_d2d(ug4b200_ctx* ctx, void* dst, const void* src, size_t bytes);
int ug4b200_memset(ug4b200_ctx* ctx, void* dst, int byte, size_t bytes);
int ug4b200_host_alloc(ug4b200_ctx* ctx, size_t bytes, void** hptr);           /* pinned */
int ug4b200_host_free(ug4b200_ctx* ctx, void* hptr);

/* ------------------------------------------------------------------ vectors
 * Vector<T> arithmetic, ugbase/lib_algebra/cpu_algebra/vector.h:124-176 and
 * ugbase/lib_algebra/common/operations_vec.h:49-175. */

int ug4b200_vec_set(ug4b200_ctx* ctx, int64_t n, double* x, double value);            /* x = value */
int ug4b200_vec_copy(ug4b200_ctx* ctx, int64_t n, double* dst, const double* src);    /* dst = src */
int ug4b200_vec_scale(ug4b200_ctx* ctx, int64_t n, double* x, double alpha);          /* x *= alpha */
int ug4b200_vec_add(ug4b200_ctx* ctx, int64_t n, double* dst, const double* src);     /* dst += src */
int ug4b200_vec_sub(ug4b200_ctx* ctx, int64_t n, double* dst, const double* src);     /* dst -= src */
/* VecScaleAdd: dest = a1*v1 + a2*v2 (+ a3*v3), evaluated left to right; dest may alias */
int ug4b200_vec_scale_add2(ug4b200_ctx* ctx, int64_t n, double* dest, double a1, const double* v1,
                           double a2, const double* v2);
int ug4b200_vec_scale_add3(ug4b200_ctx* ctx, int64_t n, double* dest, double a1, const double* v1,
                           double a2, const double* v2, double a3, const double* v3);
/* Vector::dotprod / Vector::norm (vector_impl.h:72-79, 323-329); host result, blocks on the stream */
int ug4b200_vec_dot(ug4b200_ctx* ctx, int64_t n, const double* a, const double* b, double* host_out);
int ug4b200_vec_norm(ug4b200_ctx* ctx, int64_t n, const double* a, double* host_out);
/* dst[i] = src[idx[i]]  /  dst[idx[i]] += src[i]  (surface<->level copies,
 * mg_solver_impl.hpp:211-217, 244-248); block = doubles per index */
int ug4b200_vec_gather(ug4b200_ctx* ctx, int64_t nidx, int block, double* dst, const double* src, const int* idx);
int ug4b200_vec_scatter(ug4b200_ctx* ctx, int64_t nidx, int block, double* dst, const int* idx, const double* src);
int ug4b200_vec_scatter_add(ug4b200_ctx* ctx, int64_t nidx, int block, double* dst, const int* idx, const double* src);

/* ---- device-resident scalars (Krylov loops without host round trips) ----
 * A coefficient is sign * (*dev) when dev != NULL, else the host value. */
typedef struct ug4b200_coef {
	const double* dev;  /* device pointer or NULL */
	double host;        /* value when dev == NULL; multiplier (+1/-1) when dev != NULL */
} ug4b200_coef;

int ug4b200_vec_scale_add2_ds(ug4b200_ctx* ctx, int64_t n, double* dest, ug4b200_coef a1, const double* v1,
                              ug4b200_coef a2, const double* v2);
int ug4b200_vec_scale_add3_ds(ug4b200_ctx* ctx, int64_t n, double* dest, ug4b200_coef a1, const double* v1,
                              ug4b200_coef a2, const double* v2, ug4b200_coef a3, const double* v3);

/* Device mirror of StdConvCheck (convergence_check_impl.h:85-169). */
typedef struct ug4b200_conv_state {
	double initial_defect, current_defect, last_defect;
	double min_defect, rel_reduction;
	int step, max_steps;
	int done;        /* iteration_ended() */
	int status;      /* 0 running, 1 converged, 2 max steps, 3 invalid number, 4 breakdown */
	int history_cap;
	int pad_;
	double* history; /* device array, history[k] = defect after k updates */
} ug4b200_conv_state;

/* What the last block of a device reduction does with the result r. */
enum {
	UG4B200_FIN_STORE = 0,      /* *out = r */
	UG4B200_FIN_A_DIV_R = 1,    /* *out = r; *out2 = *a / r  (CG alpha = rhoOld/lambda); r == 0 -> breakdown */
	UG4B200_FIN_R_DIV_A = 2,    /* *out2 = r / *a; *out = r  (CG beta = rho/rhoOld then rhoOld := rho when out == a) */
	UG4B200_FIN_SQRT = 3,       /* *out = sqrt(r) */
	UG4B200_FIN_CONV_START = 4, /* conv->start_defect(sqrt(r)) */
	UG4B200_FIN_CONV_UPDATE = 5 /* conv->update_defect(sqrt(r)); sets conv->done */
};
typedef struct ug4b200_fin {
	int op;
	double* out;
	double* out2;
	const double* a;
	ug4b200_conv_state* conv; /* device */
} ug4b200_fin;

int ug4b200_vec_dot_ds(ug4b200_ctx* ctx, int64_t n, const double* a, const double* b, ug4b200_fin fin);
/* dest = a1*v1 + a2*v2 and, in the same pass, sum(dest^2) -> fin  (CG: r -= alpha q; ||r||) */
int ug4b200_vec_scale_add2_norm_ds(ug4b200_ctx* ctx, int64_t n, double* dest, ug4b200_coef a1, const double* v1,
                                   ug4b200_coef a2, const double* v2, ug4b200_fin fin);
/* CG inner update in one pass: x += alpha p; r -= alpha q; sum(r^2) -> fin (cg.h:187-198) */
int ug4b200_cg_update_ds(ug4b200_ctx* ctx, int64_t n, double* x, const double* p, double* r, const double* q,
                         const double* alpha_dev, ug4b200_fin fin);
/* generic one-thread scalar program: out = (a/b)*(c/d) with NULL operands = 1 (BiCGStab beta) */
int ug4b200_scalar_ratio_ds(ug4b200_ctx* ctx, double* out, const double* a, const double* b, const double* c,
                            const double* d);
/* the same with the breakdown exits of BiCGStab (bicgstab.h:216-224 "rhoOld == 0", :374-381 "omega == 0"): a zero
 * divisor b or d ends the iteration (conv->done = 1, status 4) before anything is overwritten with inf / NaN */
int ug4b200_scalar_ratio_conv_ds(ug4b200_ctx* ctx, double* out, const double* a, const double* b, const double* c,
                                 const double* d, ug4b200_conv_state* conv);
/* apply a finaliser to a value that already sits on the device (after an all-reduce) */
int ug4b200_scalar_fin_ds(ug4b200_ctx* ctx, const double* r_dev, ug4b200_fin fin);
int ug4b200_conv_init(ug4b200_ctx* ctx, ug4b200_conv_state* dev_state, int max_steps, double min_defect,
                      double rel_reduction, double* dev_history, int history_cap);

/* ------------------------------------------------------------------ matrices
 * SparseMatrix<T> (ugbase/lib_algebra/cpu_algebra/sparsematrix.h:98, 737-747) uploaded
 * once after assembly.  Input is defragmented CRS (copy_crs, sparsematrix.h:607-617):
 * rowptr[nrows+1], cols sorted ascending inside a row, vals = block*block doubles per
 * entry, column-major inside a block.  Device layout is SELL-32 (slice = 32 rows,
 * entries stored slice-column-major so that a warp reads 32 consecutive values) with
 * true row lengths kept, explicit zeros preserved.  The handle is immutable. */

/* Scalar matrices whose values repeat (at most 65536 distinct fp64 bit patterns — the level
 * operators of a uniformly refined grid have a few dozen) and whose columns stay within 65535 of
 * the smallest column of their 32-row slice additionally get a VALUE-INDEXED copy of the entry
 * stream: u16 index into a dictionary of the exact values + u16 column offset, 4 instead of 12
 * bytes per entry (cf. CSR-VI / CSR-DU, Kourtis et al. 2008).  Lossless, so every result stays
 * bit-identical; the SpMV family streams this copy.  UG4B200_MAT_NO_COMPRESS (or the
 * environment variable UG4B200_NO_COMPRESS=1) keeps the plain stream only. */
enum { UG4B200_MAT_DEFAULT = 0, UG4B200_MAT_NO_COMPRESS = 1, UG4B200_MAT_NO_XSTAGE = 2 };

typedef struct ug4b200_matrix_info {
	int64_t nrows, ncols, nnz, padded_nnz, num_slices, device_bytes;
	int block;
	int max_row_len;
	int value_indexed;        /* 1 if the value-indexed stream exists */
	int num_distinct_values;
	int x_staged;             /* 1 if the x-staged stream exists (value-indexed words + per-slice lists of the column runs that
	                             the bulk-copy kernel stages in shared memory next to the entry words) */
	int x_staged_runs;        /* run slots per slice of that stream */
	int64_t x_staged_doubles; /* doubles of x staged per sweep (sum over the slices), aligned run ends included */
} ug4b200_matrix_info;

int ug4b200_matrix_upload_crs(ug4b200_ctx* ctx, int block, int64_t nrows, int64_t ncols, const int64_t* rowptr,
                              const int* cols, const double* vals, int flags, ug4b200_matrix** out);

/* Which entry streams ug4b200_matrix_upload_crs would build for a scalar CRS matrix, and why not — host only, no
 * device, no context (diagnosis: "why does my matrix run at 12 bytes per entry?").  Optionally returns the x-staged
 * stream itself for inspection: xw [padded_nnz] words (position << 16 | dictionary index << 3), hdr [4 * num_slices]
 * ints {entry offset / 32, width, staged bytes, runs}, runs [2 * num_slices * x_staged_runs] ints {first column,
 * doubles | position << 16}, dict [num_distinct_values] doubles; any of them may be NULL.  Call once with NULL arrays
 * to size them. */
typedef struct ug4b200_stream_plan {
	int64_t num_slices, padded_nnz;
	int max_row_len;
	int num_distinct_values;      /* -1: more than 65536 distinct values -> plain 12-byte stream */
	int64_t max_column_window;    /* largest (max - min column) of a slice; the value-indexed stream needs <= 65535 */
	int value_indexed;            /* 1: the value-indexed stream would be built */
	int x_staged;                 /* 1: the x-staged stream could be built (dictionary <= 256, rows <= 27 entries, runs fit) */
	int x_staged_runs;            /* run slots per slice it needs (limit 32) */
	int x_staged_max_doubles;     /* widest staged x segment (limit 384) */
	int64_t x_staged_doubles;     /* staged doubles summed over the slices */
} ug4b200_stream_plan;
int ug4b200_host_stream_plan(int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols, const double* vals,
                             ug4b200_stream_plan* plan, unsigned int* xw, int* hdr, int* runs, double* dict);

/* The value-indexed stream ug4b200_matrix_upload_crs would upload — host only, for inspection and tests: words
 * [padded_nnz] = (column - colbase[slice]) << 16 | dictionary index << *vshift (dictionary = the `dict` of
 * ug4b200_host_stream_plan; padding words 0), colbase [num_slices] = smallest column of the slice.  *vshift = -1 (and
 * nothing written) when the matrix gets no such stream (column window > 65535, more than 65536 distinct values, no
 * entries).  words / colbase may be NULL. */
int ug4b200_host_value_indexed_stream(int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols, const double* vals,
                                      unsigned int* words, int* colbase, int* vshift);
int ug4b200_matrix_destroy(ug4b200_ctx* ctx, ug4b200_matrix* A);
int ug4b200_matrix_get_info(const ug4b200_matrix* A, ug4b200_matrix_info* info);

/* SparseMatrix::axpy(dest, alpha, v, beta, w): dest = alpha*v + beta*A*w
 * (sparsematrix_impl.h:291-339).  alpha == 0: v ignored (may be NULL), empty rows give 0.
 * vblock = doubles per vector entry; a block-1 matrix with vblock > 1 acts per component
 * (transfer operators of a block algebra). */
int ug4b200_matrix_axpy(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* dest, double alpha, const double* v,
                        double beta, const double* w, int vblock);
/* y = A x  (SparseMatrix::apply, sparsematrix.h:184-190) */
int ug4b200_matrix_apply(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* y, const double* x, int vblock);
/* y -= A x  (SparseMatrix::matmul_minus, sparsematrix.h:199-207) */
int ug4b200_matrix_matmul_minus(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* y, const double* x, int vblock);
/* dest = beta*A*w, rows without entries untouched (apply_ignore_zero_rows, sparsematrix_impl.h:271-288) */
int ug4b200_matrix_apply_ignore_zero_rows(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* dest, double beta,
                                          const double* w, int vblock);
/* y = A x fused with the reduction (y, x) -> fin   (CG: q = A p, lambda = (q,p), cg.h:166-169) */
int ug4b200_matrix_apply_dot_ds(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* y, const double* x,
                                ug4b200_fin fin);

/* same, the reduction summed over all ranks before fin is applied (ParallelVector::dotprod of an
 * additive q with a consistent p, parallel_vector_impl.h:323-379): one kernel with the peer-window
 * transport, apply_dot -> ncclAllReduce -> finaliser through scratch_dev (1 double) otherwise */
int ug4b200_matrix_apply_dot_allreduce_ds(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* y, const double* x,
                                          ug4b200_fin fin, double* scratch_dev);

/* ------------------------------------------------------------------ smoothers */

/* Jacobi::preprocess (jacobi.h:196-220): diaginv[i] = inverse(A_ii * (1./damp));
 * block_inverse = 0 uses only the diagonal of the block (set_block(false)).
 * diaginv: device, nrows*block*block doubles (column-major blocks). */
int ug4b200_jacobi_prepare(ug4b200_ctx* ctx, const ug4b200_matrix* A, double damp, int block_inverse,
                           double* diaginv);
/* the two halves of jacobi_prepare, for partitioned runs where the additive diagonal is
 * summed over the interface copies in between (jacobi.h:171-187) */
int ug4b200_matrix_get_diag(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* diag);
int ug4b200_jacobi_invert_diag(ug4b200_ctx* ctx, int64_t nrows, int block, double damp, int block_inverse,
                               const double* diag, double* diaginv);
/* Jacobi::step (jacobi.h:222-232): c[i] = diaginv[i] * d[i] */
int ug4b200_jacobi_step(ug4b200_ctx* ctx, int64_t nblocks, int block, const double* diaginv, double* c,
                        const double* d);
/* c = diaginv*d and sc += c in one pass (first pre-smoothing step of a level) */
int ug4b200_jacobi_step_add(ug4b200_ctx* ctx, int64_t nblocks, int block, const double* diaginv, double* c,
                            const double* d, double* sc);
/* One fused smoothing step of the V-cycle (mg_solver_impl.hpp:1919-1943 with Jacobi):
 *   [sc += st_in]            if flags & UG4B200_SMOOTH_ADD_IN
 *   sd -= A*st_in
 *   [st_out = diaginv*sd]    if flags & UG4B200_SMOOTH_JACOBI      (st_out != st_in)
 *   [sc += st_out]           if flags & UG4B200_SMOOTH_ADD_OUT
 * Serial V-cycles use JACOBI|ADD_OUT; partitioned ones ADD_IN|JACOBI because st_out must be
 * made consistent across ranks before it is accumulated. */
/* UG4B200_SMOOTH_SC_ZERO: the incoming sc is taken as 0.0 without reading it (the level's
 * correction was just reset, mg_solver_impl.hpp:234, 1780) — saves the set(0) sweep; the
 * additions 0.0 + st are still carried out, so sc is bit-identical. */
enum { UG4B200_SMOOTH_ADD_IN = 1, UG4B200_SMOOTH_JACOBI = 2, UG4B200_SMOOTH_ADD_OUT = 4, UG4B200_SMOOTH_SC_ZERO = 8 };
int ug4b200_jacobi_smooth_fused(ug4b200_ctx* ctx, const ug4b200_matrix* A, const double* diaginv, double* sd,
                                const double* st_in, double* st_out, double* sc, int flags);
/* same with the incoming defect read from sd_in (sd = 1.0*sd_in - A*st_in): the top level
 * smooths the caller's defect without the surface->level copy of mg_solver_impl.hpp:211-217
 * when the index map is the identity.  sd_in == NULL or == sd: in place. */
int ug4b200_jacobi_smooth_fused_src(ug4b200_ctx* ctx, const ug4b200_matrix* A, const double* diaginv, double* sd,
                                    const double* sd_in, const double* st_in, double* st_out, double* sc, int flags);
/* Restriction fused with the first Jacobi step of the coarse level
 * (mg_solver_impl.hpp:1802 do_restrict -> std_transfer_impl.h:791-792, then :1705 on the
 * coarse level): sd_coarse = beta*R*sd_fine, rows without connections untouched;
 * st_coarse = diaginv_coarse * sd_coarse.  Scalar algebra. */
int ug4b200_restrict_jacobi_fused(ug4b200_ctx* ctx, const ug4b200_matrix* R, const double* diaginv_coarse,
                                  double* sd_coarse, double beta, const double* sd_fine, double* st_coarse);

/* Multicolour Gauss-Seidel.  The matrix must be given in a colour-sorted DoF order:
 * rows [color_ptr[k], color_ptr[k+1]) form colour k and have no stored connection to
 * another row of the same colour.  A lexicographic gs_step_LL / gs_step_UR
 * (ugbase/lib_algebra/algebra_common/core_smoothers.h:105-169) over such a matrix IS
 * multicolour GS, so results are bit-identical to the CPU sweeps on the same matrix. */
int ug4b200_color_greedy(int64_t nrows, const int64_t* rowptr, const int* cols, int* color, int* ncolors); /* host */
int ug4b200_color_check(int64_t nrows, const int64_t* rowptr, const int* cols, int ncolors,
                        const int64_t* color_ptr); /* host; 0 if the order is a valid colouring */
/* kind: 0 forward (gs_step_LL), 1 backward (gs_step_UR), 2 symmetric (sgs_step) */
int ug4b200_gs_step(ug4b200_ctx* ctx, const ug4b200_matrix* A, int ncolors, const int64_t* color_ptr_host,
                    int kind, double relax, double* c, const double* d);

/* dense LU base solver (lu.h:122-140, 189-207): factors computed on the host by the
 * caller (row-major n*n LU with unit lower part, pivot interchanges as in SolveLU,
 * no_lapack/lu_decomp.h:160-195), applied on the device by one CTA. */
int ug4b200_lu_apply(ug4b200_ctx* ctx, int n, const double* lu_dev, const int* piv_dev, double* x, const double* b);
/* small on-device CG base solver: one CTA iterates x (start 0) on A x = b until
 * ||r|| < max(min_defect, rel_reduction*||r0||) or max_steps (north_star: "coarse-grid solve
 * done as a small on-device CG") */
int ug4b200_coarse_cg(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* x, const double* b, double* work4n,
                      int max_steps, double min_defect, double rel_reduction);

/* ------------------------------------------------------------------ multi-GPU
 * pcl::InterfaceCommunicator + ProcessCommunicator semantics over NCCL
 * (ugbase/pcl/pcl_interface_communicator_impl.hpp:408-739,
 *  ugbase/pcl/pcl_process_communicator.cpp:311-326). */

#define UG4B200_NCCL_ID_BYTES 128
int ug4b200_comm_unique_id(unsigned char id[UG4B200_NCCL_ID_BYTES]);
int ug4b200_comm_init(ug4b200_ctx* ctx, int nranks, int rank, const unsigned char id[UG4B200_NCCL_ID_BYTES]);
int ug4b200_comm_destroy(ug4b200_ctx* ctx);
/* in-place sum over ranks of n device doubles (ParallelVector::dotprod/norm allreduce,
 * parallel_vector_impl.h:269-379) */
int ug4b200_allreduce_sum(ug4b200_ctx* ctx, double* dev, int n);
/* Horizontal interface: for every neighbour rank p the list of local block indices shared
 * with p, in an order both sides agree on (IndexLayout, parallel_index_layout.h:52-53).
 * owner[i] = 1 if this rank is the h-master (lowest rank) of interface index list entry. */
int ug4b200_interface_create(ug4b200_ctx* ctx, int nneigh, const int* neigh_rank, const int64_t* neigh_ptr,
                             const int* indices, int64_t nlocal, ug4b200_interface** out);
int ug4b200_interface_destroy(ug4b200_ctx* ctx, ug4b200_interface* I);
/* AdditiveToConsistent (parallelization_util.h:159-191): every copy of an interface
 * DoF ends up with the sum over all copies, summed in ascending rank order on every
 * rank (bitwise identical copies). */
int ug4b200_additive_to_consistent(ug4b200_ctx* ctx, ug4b200_interface* I, double* v, int block);
/* AdditiveToUnique (parallelization_util.h:260-280): the h-master copy gets the sum over all
 * copies (same ascending-rank order), every other copy becomes 0 — one kernel with the
 * peer-window transport */
int ug4b200_additive_to_unique(ug4b200_ctx* ctx, ug4b200_interface* I, double* v, int block);
/* Fused interface push (peer-window transport, scalar vectors): announce that the NEXT smoothing
 * kernel whose output vector is `vec` (ug4b200_jacobi_smooth_fused* / ug4b200_restrict_jacobi_fused)
 * is followed by ug4b200_additive_to_consistent(I, vec).  That kernel then stores the interface rows
 * of its result straight into the neighbours' peer windows while it streams the matrix and its last
 * CTA raises the flags, so the exchange itself only waits for the neighbours and adds the copies
 * (same ascending-rank sum, bit-identical).  If the producing call cannot push (recorded small
 * operation, block vectors, NCCL transport) the arm is dropped and the exchange runs in full.
 * Opt-in (environment UG4B200_FUSED_PUSH=1): at N = 2 it measured slower than the one-kernel
 * exchange; without it this call is a no-op. */
int ug4b200_interface_arm(ug4b200_ctx* ctx, ug4b200_interface* I, const double* vec);
/* Peer-window transport (preferred inside one NVSwitch box; replaces the MPI_Isend/Irecv
 * transport of pcl_interface_communicator_impl.hpp:560-661 and MPI_Allreduce,
 * pcl_process_communicator.cpp:325): every rank exposes a window of device memory to the
 * other ranks of the box (CUDA IPC); interface exchange and scalar all-reduce then run as
 * ONE kernel per rank that stores straight into the neighbours' windows over NVLink and
 * synchronises through epoch flags there — no NCCL call, no host round trip.
 *   1. window_create on every rank (bytes == 0: UG4B200_P2P_WINDOW_MB or 64 MiB),
 *   2. exchange the 64-byte handles (the caller's plumbing, e.g. torch.distributed),
 *   3. window_open with all handles in rank order (or window_attach with raw base
 *      pointers when the "ranks" are contexts of one process),
 *   4. create interfaces: all ranks must create (and destroy) them in the same order.
 * Interfaces created while a window is open use it; ug4b200_allreduce_sum does for
 * n <= 1024.  Results are identical to the NCCL transport (same ascending-rank sums). */
#define UG4B200_IPC_HANDLE_BYTES 64
int ug4b200_p2p_window_create(ug4b200_ctx* ctx, size_t bytes, unsigned char handle[UG4B200_IPC_HANDLE_BYTES],
                              void** base);
int ug4b200_p2p_window_open(ug4b200_ctx* ctx, int nranks, int rank, const unsigned char* handles);
int ug4b200_p2p_window_attach(ug4b200_ctx* ctx, int nranks, int rank, void* const* bases);
int ug4b200_p2p_window_destroy(ug4b200_ctx* ctx);
int ug4b200_p2p_enabled(const ug4b200_ctx* ctx);
/* non-zero if a kernel gave up waiting for a neighbour (20 s); also reported by ug4b200_sync */
int ug4b200_p2p_check(ug4b200_ctx* ctx);
/* resolve where the neighbours receive (blocks until they have created the same interface);
 * done implicitly by the first exchange, but must happen before a graph capture */
int ug4b200_interface_commit(ug4b200_ctx* ctx, ug4b200_interface* I);
/* ParallelVector::dotprod / norm (parallel_vector_impl.h:269-379) without leaving the device:
 * local dot, sum over ranks and finaliser.  One kernel with the peer-window transport;
 * dot -> ncclAllReduce -> finaliser through scratch_dev (1 double) otherwise. */
int ug4b200_vec_dot_allreduce_ds(ug4b200_ctx* ctx, int64_t n, const double* a, const double* b, ug4b200_fin fin,
                                 double* scratch_dev);
/* Gathered level (mg_solver_impl.hpp:2003-2070: gather the defect of a coarse level, solve
 * there, hand the correction back).  Every rank holds the whole coarse level; gather_sum turns
 * the local ADDITIVE vectors into the global vector summed over all ranks (ascending rank
 * order, identical on every rank) — one kernel over the peer windows, or set/scatter/
 * ncclAllReduce.  local_to_global: global block index of every local block index. */
typedef struct ug4b200_gather ug4b200_gather;
int ug4b200_gather_create(ug4b200_ctx* ctx, int64_t nglobal, int64_t nlocal, const int* local_to_global, int block,
                          ug4b200_gather** out);
int ug4b200_gather_commit(ug4b200_ctx* ctx, ug4b200_gather* G);
int ug4b200_gather_sum(ug4b200_ctx* ctx, ug4b200_gather* G, double* global_out, const double* local_in);
int ug4b200_gather_destroy(ug4b200_ctx* ctx, ug4b200_gather* G);
/* zero every copy that is not the h-master (AdditiveToUnique after a consistent sum /
 * ConsistentToUnique, parallelization_util.h:260-280, 387-393) */
int ug4b200_set_slaves_zero(ug4b200_ctx* ctx, ug4b200_interface* I, double* v, int block);
/* sum over h-master/inner entries of a[i]*b[i] (unique dot of consistent vectors) */
int ug4b200_vec_dot_unique_ds(ug4b200_ctx* ctx, ug4b200_interface* I, int64_t n, int block, const double* a,
                              const double* b, double* out_dev);

#ifdef __cplusplus
}
#endif
#endif /* UG4B200_H */
