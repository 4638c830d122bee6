/*
 * ug4b200_solver.h — descriptor-level C ABI on top of the host-side mirror of ugcore's
 * operator API (ugcore_b200/csrc/host/: GPUVector, GPUSparseMatrix, CG, BiCGStab,
 * LinearSolver, Jacobi, GaussSeidel*, LU, StdTransfer, AssembledMultiGridCycle).
 *
 * It plays the role of scripts/util/solver_util.lua (util.solver.CreateSolver, :602;
 * SolveLinearProblem, :1182-1210) for callers that are not Lua: build the solver objects
 * from a descriptor, hand over the assembled level matrices, then solver:init(A,u) and
 * solver:apply(u,b) (bound in ugbase/bridge/algebra_bridges/common_bridge.cpp:326-338).
 * Used by the Python package, the tests and bench.py.  Everything here runs on the GPU;
 * there is no CPU fallback.
 */
#ifndef UG4B200_SOLVER_H
#define UG4B200_SOLVER_H
#include "ug4b200.h"
#ifdef __cplusplus
extern "C" {
#endif

enum { UG4B200_SOLVER_CG = 0, UG4B200_SOLVER_BICGSTAB = 1, UG4B200_SOLVER_LINEAR = 2, UG4B200_SOLVER_LU = 3,
       UG4B200_SOLVER_COARSE_CG = 4, UG4B200_SOLVER_GMRES = 5 /* lib_algebra/operator/linear_solver/gmres.h */ };
enum { UG4B200_PRECOND_NONE = 0, UG4B200_PRECOND_JACOBI = 1, UG4B200_PRECOND_GS = 2, UG4B200_PRECOND_BGS = 3,
       UG4B200_PRECOND_SGS = 4, UG4B200_PRECOND_GMG = 5, UG4B200_PRECOND_ILU = 6 /* operator/preconditioner/ilu.h */ };
/* ordering of the ILU factorisation (ILU::set_ordering_algorithm / set_sort, ilu.h:397-414) */
enum { UG4B200_ILU_ORDER_NATURAL = 0,    /* no ordering: level-scheduled triangular solves, one launch per level */
       UG4B200_ILU_ORDER_CMK = 1,        /* set_sort(true): NativeCuthillMcKeeOrdering; level-scheduled */
       UG4B200_ILU_ORDER_MULTICOLOR = 2  /* greedy multicolour ordering: one launch per colour, bit-identical to the
                                            reference's ILU applied in that ordering */ };
enum { UG4B200_FLAG_HOST_SCALARS = 1,   /* CG / BiCGStab / LinearSolver as reference-shaped host loops (one sync per dot / norm) instead of
                                           the default: scalars and convergence state on the device, one CUDA graph per iteration */
       UG4B200_FLAG_NO_GRAPH = 2,       /* do not capture the Krylov iteration into a CUDA graph */
       UG4B200_FLAG_NO_FUSED_JACOBI = 4,/* V-cycle with separate Jacobi / SpMV / AXPY launches */
       UG4B200_FLAG_FINAL_LEVEL_DEFECT = 8, /* also do the reference's unused top-level defect update */
       UG4B200_FLAG_DEVICE_LINEAR = 64,   /* accepted, no effect: the device-resident LinearSolver is the default since round 2 */
       UG4B200_FLAG_DEVICE_BICGSTAB = 32, /* accepted, no effect: the device-resident BiCGStab is the default since round 2
                                             (measured 33.3 vs 38.6 ms per solve on configs[3] at 129^3, identical histories) */
       UG4B200_FLAG_RAP = 16            /* gmg:set_rap(true): level operators below the top level are the Galerkin
                                           products R A P (mg_solver_impl.hpp:828-1013), computed on the host at init;
                                           ug4b200_solver_set_level then takes rowptr == NULL on those levels */ };

typedef struct ug4b200_solver_desc {
	int block;             /* 1 = GPUAlgebra, 2/3 = GPUBlockAlgebra<N> */
	int solver;            /* UG4B200_SOLVER_* */
	int precond;           /* UG4B200_PRECOND_* */
	double damp;           /* Jacobi damping / GS relaxation when used directly as preconditioner */
	int max_steps;         /* StdConvCheck(maxSteps, minDefect, relReduction) */
	double min_defect;
	double rel_reduction;
	/* GeometricMultiGrid (solver_util.lua:466-483) */
	int base_lev, top_lev;
	int cycle;             /* 1 V, 2 W, -1 F */
	int nu1, nu2;
	int smoother;          /* UG4B200_PRECOND_JACOBI / GS / BGS / SGS */
	double smoother_damp;
	int base_solver;       /* UG4B200_SOLVER_LU or UG4B200_SOLVER_COARSE_CG */
	int base_max_steps;
	double base_min_defect, base_rel_reduction;
	int flags;             /* UG4B200_FLAG_* */
	/* partitioned runs: levels base_lev..gather_lev are gathered — every rank holds them completely
	 * and runs that part of the V-cycle redundantly (set_gathered_base / set_gathered_level), the
	 * levels above are partitioned.  gather_lev <= base_lev: only the base solve is gathered
	 * (mg_solver_impl.hpp:2003-2070). */
	int gather_lev;
	int restart;           /* GMRES(restart); <= 0: 5 (solver_util.lua:669-670 creates GMRES(5)).  BiCGStab: set_restart(restart),
	                          <= 0: no periodic restart (bicgstab.h:161-163) */
	double ilu_beta;       /* ILU(beta), 0 = ILU(0); for precond == ILU and for smoother == ILU inside GMG */
	int ilu_order;         /* UG4B200_ILU_ORDER_* */
} ug4b200_solver_desc;

typedef struct ug4b200_solver ug4b200_solver;

/* process-wide device context of the host layer (GPUManager); device < 0: LOCAL_RANK */
int ug4b200_host_init(int device, void* stream);
int ug4b200_host_finalize(void);
ug4b200_ctx* ug4b200_host_ctx(void);
const char* ug4b200_host_last_error(void);
/* partitioned runs: NCCL communicator of the host layer's context */
int ug4b200_host_comm_init(int nranks, int rank, const unsigned char id[UG4B200_NCCL_ID_BYTES]);

int ug4b200_solver_create(const ug4b200_solver_desc* d, ug4b200_solver** out);
int ug4b200_solver_destroy(ug4b200_solver* s);
/* the assembled surface matrix A (AssembledLinearOperator); host CRS, copied */
int ug4b200_solver_set_matrix(ug4b200_solver* s, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols,
                              const double* vals);
/* GMG level lev: assembled level matrix and the transfers P (lev-1 -> lev), R (lev -> lev-1);
 * P/R are scalar CRS and NULL on the base level; the top level may reuse the surface matrix
 * (pass rowptr == NULL). */
int ug4b200_solver_set_level(ug4b200_solver* s, int lev, int64_t nrows, const int64_t* rowptr, const int* cols,
                             const double* vals, int64_t ncoarse, const int64_t* p_rowptr, const int* p_cols,
                             const double* p_vals, const int64_t* r_rowptr, const int* r_cols, const double* r_vals);
/* surface <-> level index map of the GMG (vSurfLevelMap, mg_solver_impl.hpp:1344-1369): surface index of every top-level
 * index, for hierarchies whose surface DoF distribution numbers the top level differently from the level DoF
 * distribution (the matrix given to set_matrix and the vectors of apply are in SURFACE numbering, the level matrices
 * and transfers in LEVEL numbering).  Without it the map is the identity and the surface <-> level copies of
 * AssembledMultiGridCycle::apply (:211-217, 244-248) are elided. */
int ug4b200_solver_set_surface_map(ug4b200_solver* s, int64_t n, const int* surf_index_of_level_index);
/* colour-sorted DoF order for the Gauss-Seidel smoother of level lev (lev = -1: the direct
 * preconditioner): perm maps old -> new index, colours are [color_ptr[k], color_ptr[k+1]) */
int ug4b200_solver_set_coloring(ug4b200_solver* s, int lev, int64_t n, const int* perm, int ncolors,
                                const int64_t* color_ptr);
/* horizontal interfaces of level lev (lev = top level also serves the Krylov vectors) */
int ug4b200_solver_set_layouts(ug4b200_solver* s, int lev, int nneigh, const int* neigh_rank, const int64_t* neigh_ptr,
                               const int* indices, int64_t nlocal);
/* partitioned Gauss-Seidel smoothing on level lev (the top level also serves a Gauss-Seidel used directly as
 * preconditioner): the level matrix made CONSISTENT on the interface rows, i.e. every copy of an
 * interface row holds the sum over the ranks of the entries whose two DoFs the rank holds — what
 * GaussSeidelBase::preprocess obtains from MakeConsistent(*pOp, m_A)
 * (ugbase/lib_algebra/operator/preconditioner/gauss_seidel.h:134-142,
 * ugbase/lib_algebra/parallelization/parallel_matrix_overlap_impl.h:438-459).  ugcore exchanges the rows
 * over MPI inside preprocess; here the caller exchanges them on the host at init
 * (ugcore_b200/dist.py: make_consistent).  The rows of h-slaves are set to Dirichlet rows by the
 * smoother itself.  Same pattern as the additive level matrix. */
int ug4b200_solver_set_smoother_matrix(ug4b200_solver* s, int lev, int64_t nrows, const int64_t* rowptr, const int* cols,
                                       const double* vals);
/* gathered base solve: global base-level matrix and the local -> global index map */
int ug4b200_solver_set_gathered_base(ug4b200_solver* s, int64_t nrows, const int64_t* rowptr, const int* cols,
                                     const double* vals, int64_t nlocal, const int* local_to_global);
/* gathered cycle (desc.gather_lev > base_lev): GLOBAL level matrix and transfers of level lev,
 * base_lev <= lev <= gather_lev; the matrix of gather_lev itself is the one given to
 * set_gathered_base (pass rowptr == NULL here), P/R are NULL on base_lev */
int ug4b200_solver_set_gathered_level(ug4b200_solver* s, int lev, int64_t nrows, const int64_t* rowptr, const int* cols,
                                      const double* vals, int64_t ncoarse, const int64_t* p_rowptr, const int* p_cols,
                                      const double* p_vals, const int64_t* r_rowptr, const int* r_cols, const double* r_vals);
/* solver:set_debug(GridFunctionDebugWriter): while CG runs, the residual and the solution are written after every step
 * as <dir>/CG_Residual_iterNNN.vec and <dir>/CG_Solution_iterNNN.vec in ConnectionViewer form (write_debugXR,
 * ugbase/lib_algebra/operator/linear_solver/cg.h:124, 195, 273-280) — the files a ugcore run with a debug writer leaves,
 * for side-by-side comparison.  positions: 3 doubles per node or NULL (zeros); precision 0 = the
 * reference's 16 digits, 17 = lossless.  The solve then runs the host-paced loop (one sync per step).  dir == NULL
 * switches the writer off again. */
int ug4b200_solver_set_debug_dir(ug4b200_solver* s, const char* dir, const double* positions, int64_t npos, int dim, int precision);
/* solver:init(A, u): uploads, smoother preprocess, base factorisation */
int ug4b200_solver_init(ug4b200_solver* s);
/* solver:apply(u, b) with HOST vectors: H2D of x and b, solve, D2H of x.
 * Returns 0 converged, 1 not converged / breakdown, < 0 error. */
int ug4b200_solver_apply(ug4b200_solver* s, double* x_host, const double* b_host);
/* the same for the usual call with a zero start vector (u:set(0.0) before solver:apply, solver_util.lua:1194-1200):
 * x is set to 0 on the device instead of being copied there — a third less host-to-device traffic; x_host is output only */
int ug4b200_solver_apply_zero_guess(ug4b200_solver* s, double* x_host, const double* b_host);
/* same with DEVICE vectors (n = block * rows doubles) */
int ug4b200_solver_apply_device(ug4b200_solver* s, double* x_dev, const double* b_dev);
int ug4b200_solver_steps(const ug4b200_solver* s);
double ug4b200_solver_defect(const ug4b200_solver* s);
/* defect history, entry 0 = start defect; returns the number of entries copied */
int ug4b200_solver_history(const ug4b200_solver* s, double* out, int cap);
/* c = M^-1 d with the configured preconditioner alone (host vectors) */
int ug4b200_solver_precond_apply(ug4b200_solver* s, double* c_host, const double* d_host);
int64_t ug4b200_solver_num_dofs(const ug4b200_solver* s);

/* ---- import / export of assembled objects (host side, no device involved) -------------------
 * ConnectionViewer .mat / .vec (ugbase/lib_algebra/common/connection_viewer_{output,input}.h) and
 * MatrixMarket .mtx (ugbase/lib_algebra/common/matrixio/matrix_io_mtx.{h,cpp}): the formats in which
 * a ugcore installation dumps A, the level matrices, P, R, b and per-iteration residuals (debug
 * writers, cg.h:274-280, mg_solver_impl.hpp:692-696, 2181-2200).  Scalar matrices.
 * format: 0 = by file extension (.mtx -> MatrixMarket, anything else ConnectionViewer), 1 = ConnectionViewer,
 * 2 = MatrixMarket. */
typedef struct ug4b200_host_matrix ug4b200_host_matrix;
/* keep_zeros = 0 follows the reference reader (zero values are not inserted).  n_to > 0: a ConnectionViewer
 * file written with from / to positions (rectangular P / R): rows 0..n_to-1, columns numbered behind them. */
int ug4b200_io_read_matrix(const char* filename, int format, int keep_zeros, int64_t n_to, ug4b200_host_matrix** out);
int ug4b200_io_matrix_info(const ug4b200_host_matrix* m, int64_t* nrows, int64_t* ncols, int64_t* nnz, int* dim, int64_t* npos);
/* rowptr[nrows+1], cols[nnz], vals[nnz]; positions[3*npos] or NULL */
int ug4b200_io_matrix_export(const ug4b200_host_matrix* m, int64_t* rowptr, int* cols, double* vals, double* positions);
void ug4b200_io_matrix_free(ug4b200_host_matrix* m);
/* positions: 3 doubles per position (NULL: all zero).  ConnectionViewer: npos = nrows positions, or — from / to
 * form, from_to = 1 — nrows + ncols positions (rows first).  precision 0 = the reference's formatting
 * (ConnectionViewer: 6 significant digits for matrix values; MatrixMarket: 13 digits), > 0 = that many
 * digits (17 resp. 16 are lossless for fp64). */
int ug4b200_io_write_matrix(const char* filename, int format, int64_t nrows, int64_t ncols, const int64_t* rowptr,
                            const int* cols, const double* vals, const double* positions, int dim, int from_to,
                            int precision);
int ug4b200_io_vector_size(const char* filename, int64_t* n, int* dim);
int ug4b200_io_read_vector(const char* filename, int64_t n, double* values, double* positions);
/* M = R * A * P by the reference's AddMultiplyOf loop (sparsematrix_util.h:152-230; bit-identical): scalar
 * CRS in (R: nc x nf, A: nf x nf, P: nf x nc), host matrix handle out (read with ug4b200_io_matrix_info /
 * _export, release with _free).  The solver does this itself under UG4B200_FLAG_RAP (also for block
 * matrices); the entry point exists for callers that want the coarse operator alone. */
int ug4b200_host_rap(int64_t nc, int64_t nf, const int64_t* r_rowptr, const int* r_cols, const double* r_vals,
                     const int64_t* a_rowptr, const int* a_cols, const double* a_vals,
                     const int64_t* p_rowptr, const int* p_cols, const double* p_vals, ug4b200_host_matrix** out);
int ug4b200_io_write_vector(const char* filename, int64_t n, const double* values, const double* positions, int dim,
                            int precision);

/* ---- assembly-side API of the GPU matrix type (what DomainDiscretization, constraints and transfers call on
 * matrix_type before the solve; ugbase/lib_algebra/cpu_algebra/sparsematrix.h:116-343): a small
 * interpreter so that callers without C++ (and the tests) can drive it.  Scalar matrices.  ops: nops x 4 doubles
 * (code, r, c, v):
 *   0 resize_and_clear(r, c)        1 A(r,c) = v                   2 A(r,c) += v             3 scale(v)
 *   4 clear_retain_structure()      5 resize_and_keep_values(r, c) 6 defragment()            7 set(v): diagonal v, rest 0
 *   8 A := transpose(A) * v         9 A := copy of A * v           10 read A(r,c) through the const access (creates nothing)
 *  11 set_matrix_row(r): the following c ops of code 12 (.., col, value) form the row
 *  13 add_matrix_row(r): likewise
 * Returns a host matrix handle (ug4b200_io_matrix_info / _export / _free). */
int ug4b200_host_matrix_script(int64_t nops, const double* ops, ug4b200_host_matrix** out);
/* is_isolated(i) for every row of a host matrix (sparsematrix_impl.h:416-425) */
int ug4b200_host_matrix_isolated(const ug4b200_host_matrix* m, unsigned char* isolated);

/* y = A^T x on the device through the matrix type's apply_transposed (sparsematrix.h:194-197, sparsematrix_impl.h:341-370):
 * host CRS in (block x block entries), host vectors in / out; for callers and tests — the solve path itself applies
 * explicit restriction matrices and never needs the transposed product (std_transfer_impl.h:694-695). */
int ug4b200_host_apply_transposed(int block, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols,
                                  const double* vals, double* y_host, const double* x_host);

/* assembly-side API of the GPU vector type (ugbase/lib_algebra/cpu_algebra/vector.h:63-221), for callers and tests:
 * srand(seed); v.set_random(from, to)  -> values_out[n*block]  (same numbers as ugcore's Vector for the same seed);
 * then v.add(add_vals, idx, nidx) on the host mirror, one device operation (v *= 2) and the read-back through
 * get(idx) -> got_out[nidx*block]; maxnorm_out = v.maxnorm().  Exercises the host mirror <-> device hand-over. */
int ug4b200_host_vector_selftest(int block, int64_t n, unsigned seed, double from, double to, int64_t nidx, const int64_t* idx,
                                 const double* add_vals, double* values_out, double* got_out, double* maxnorm_out);

/* ---- init-time host kernels of ILU and of DoF reordering, exposed for callers and tests (no device involved) ----
 * ILU(0) (beta == 0: FactorizeILUSorted, ilu.h:174-228) or ILU(beta) (FactorizeILUBeta, :110-171) of a scalar CRS
 * matrix with sorted rows, in place in vals: L below the diagonal (unit diagonal implied), U on and above it.
 * Bit-identical to the reference's factorisation. */
int ug4b200_host_ilu_factorize(int64_t n, const int64_t* rowptr, const int* cols, double* vals, double beta, double sort_eps);
/* same for block x block entries (block*block doubles per entry, column-major; block = 1, 2, 3) */
int ug4b200_host_ilu_factorize_block(int block, int64_t n, const int64_t* rowptr, const int* cols, double* vals, double beta,
                                     double sort_eps);
/* level sets of the lower (lower != 0) or upper triangle of a stored pattern: level[i] = longest dependency chain below
 * row i — rows of one level can be solved in one launch of the sweep kernel; *nlevels out */
int ug4b200_host_level_sets(int64_t n, const int64_t* rowptr, const int* cols, int lower, int* level, int* nlevels);
/* Cuthill-McKee ordering of the matrix graph, new_index[old] = new: GetCuthillMcKeeOrder
 * (ugbase/lib_algebra/algebra_common/permutation_util.h:96-114) -> ComputeCuthillMcKeeOrder
 * (ugbase/lib_algebra/ordering_strategies/algorithms/native_cuthill_mckee.cpp:100-300); same result as the
 * reference for the same pattern (stable sorts by degree, breadth-first numbering). */
int ug4b200_host_cuthill_mckee(int64_t n, const int64_t* rowptr, const int* cols, int reverse, int preserve_consec,
                               int64_t* new_index);

#ifdef __cplusplus
}
#endif
#endif
