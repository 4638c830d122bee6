/*
 * oracle/oracle.h — C ABI of the CPU oracle.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs only.  The product (ugcore_b200) never links or calls this.
 *
 * Two builds export the same symbols:
 *   oracle/liboracle.so            backend "port"  (restated arithmetic)
 *   oracle/_ref/liboracle_ref.so   backend "ref"   (real ugcore templates compiled
 *                                  from /root/reference/ugbase; solver control
 *                                  flow still restated, see solvers.h)
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_mat oracle_mat;

const char* oracle_backend_name(void);
const char* oracle_last_error(void);

/* ---- kernel level (raw host arrays; vectors hold block*n doubles) ---- */
oracle_mat* oracle_mat_create(int block, int64_t nrows, int64_t ncols, const int64_t* rowptr,
                              const int* cols, const double* vals);
void    oracle_mat_destroy(oracle_mat* A);
int64_t oracle_mat_nnz(const oracle_mat* A);
int64_t oracle_mat_rows(const oracle_mat* A);
int64_t oracle_mat_cols(const oracle_mat* A);
int     oracle_mat_export(const oracle_mat* A, int64_t* rowptr, int* cols, double* vals);
oracle_mat* oracle_mat_transpose(const oracle_mat* A, int keep_zeros);
/* R * A * P by AddMultiplyOf (sparsematrix_util.h:152-230), the Galerkin coarse operator of gmg:set_rap(true) */
oracle_mat* oracle_mat_rap(const oracle_mat* R, const oracle_mat* A, const oracle_mat* P);

/* dest = alpha*v + beta*A*w (SparseMatrix::axpy); v may be NULL when alpha == 0;
 * v == dest selects the in-place branch */
int oracle_axpy(const oracle_mat* A, double* dest, double alpha, const double* v, double beta, const double* w, int vblock);
int oracle_apply(const oracle_mat* A, double* y, const double* x, int vblock);
/* srand(seed); Vector::set_random(from, to) (vector_impl.h:91-96) and its maxnorm (:332-338) */
int oracle_vec_set_random(int64_t nblocks, int block, unsigned seed, double from, double to, double* out, double* maxnorm);
/* y = A^T x (SparseMatrix::apply_transposed -> axpy_transposed, sparsematrix_impl.h:341-370) */
int oracle_apply_transposed(const oracle_mat* A, double* y, const double* x);
int oracle_matmul_minus(const oracle_mat* A, double* y, const double* x, int vblock);
int oracle_apply_ignore_zero_rows(const oracle_mat* A, double* dest, double beta, const double* w, int vblock);

double oracle_dot(int64_t nblocks, int block, const double* a, const double* b);
double oracle_norm(int64_t nblocks, int block, const double* a);
int oracle_scale_add2(int64_t len, double* d, double a1, const double* v1, double a2, const double* v2);
int oracle_scale_add3(int64_t len, double* d, double a1, const double* v1, double a2, const double* v2, double a3, const double* v3);

/* c = Jacobi(A, damp) d */
int oracle_jacobi(const oracle_mat* A, double damp, int block_inverse, double* c, const double* d);
/* kind: 0 gs_step_LL, 1 gs_step_UR, 2 sgs_step */
int oracle_gs(const oracle_mat* A, int kind, double relax, double* c, const double* d);
/* ILU(0) / ILU(beta) of a copy of A (lib_algebra/operator/preconditioner/ilu.h:110-228) and one application
 * c = U^-1 L^-1 d (invert_L :233-252, invert_U :257-322; returns 1 if the last row's near-zero check fired) */
/* SparseMatrix's assembly-side API driven by an op script (codes: include/ug4b200_solver.h, ug4b200_host_matrix_script);
 * ref backend only */
oracle_mat* oracle_mat_script(int64_t nops, const double* ops, unsigned char* isolated, int64_t isolated_cap);
oracle_mat* oracle_ilu_factorize(const oracle_mat* A, double beta, double sort_eps);
int oracle_ilu_apply(const oracle_mat* LU, double inv_eps, double* c, const double* d);
/* new_index[old] = new; GetCuthillMcKeeOrder (algebra_common/permutation_util.h:96-114) */
int oracle_cuthill_mckee(const oracle_mat* A, int reverse, int preserve_consec, int64_t* new_index);
int oracle_lu_solve(const oracle_mat* A, double* x, const double* b);

/* ---- solver level ---- */
enum { ORACLE_SOLVER_CG = 0, ORACLE_SOLVER_BICGSTAB = 1, ORACLE_SOLVER_LINEAR = 2, ORACLE_SOLVER_LU = 3 };
enum { ORACLE_PRECOND_NONE = 0, ORACLE_PRECOND_JACOBI = 1, ORACLE_PRECOND_GS = 2, ORACLE_PRECOND_BGS = 3,
       ORACLE_PRECOND_SGS = 4, ORACLE_PRECOND_GMG = 5, ORACLE_PRECOND_ILU = 6 };
enum { ORACLE_SOLVER_GMRES = 5 };

typedef struct oracle_solver_desc {
	int solver;            /* ORACLE_SOLVER_* */
	int precond;           /* ORACLE_PRECOND_* */
	double damp;           /* Jacobi damping / GS relax when used directly as preconditioner */
	int max_steps;         /* StdConvCheck */
	double min_defect;
	double rel_reduction;
	/* GMG */
	int base_lev, top_lev;
	int cycle;             /* 1 V, 2 W, -1 F */
	int nu1, nu2;
	int smoother;          /* ORACLE_PRECOND_JACOBI/GS/BGS/SGS */
	double smoother_damp;  /* Jacobi: damping; GS family: relax */
	int base_solver;       /* ORACLE_SOLVER_LU or ORACLE_SOLVER_CG (unpreconditioned, tight tolerance) */
	int base_max_steps;
	double base_min_defect, base_rel_reduction;
	int restart;           /* GMRES(restart); BiCGStab: numRestarts (0 = never) */
	double ilu_beta;       /* ILU(beta); 0: ILU(0) (also for an ILU smoother inside GMG: smoother = ORACLE_PRECOND_ILU) */
} oracle_solver_desc;

typedef struct oracle_solver oracle_solver;

oracle_solver* oracle_solver_create(const oracle_solver_desc* d);
void oracle_solver_destroy(oracle_solver* s);
/* 0 (default): the reference's sequential dot / norm; 1: pairwise-tree summation of the same products — NOT ugcore's
 * arithmetic, a measurement aid: how far does the reference's own history move under another summation order? */
void oracle_set_reduction_mode(int mode);
int oracle_reduction_mode(void);
/* GMG level operators; P, R may be NULL on the base level. Matrices stay owned by the caller. */
int oracle_solver_set_level(oracle_solver* s, int lev, const oracle_mat* A, const oracle_mat* P, const oracle_mat* R);
/* matrix the smoothers of GMG level lev are initialised with instead of the level operator (ugcore's
 * parallel Gauss-Seidel smooths with its own consistent matrix, gauss_seidel.h:134-142) */
int oracle_solver_set_smoother_matrix(oracle_solver* s, int lev, const oracle_mat* S);
/* one-level preconditioner (Jacobi / GS / ILU) initialised with M instead of the solver's matrix: ugcore's parallel
 * Gauss-Seidel / ILU precondition with their own consistent matrix m_A / m_ILU (gauss_seidel.h:134-142, ilu.h:536-543) */
int oracle_solver_set_precond_matrix(oracle_solver* s, const oracle_mat* M);
/* surface index of every top-level index (GMG::apply's surface <-> level copies, mg_solver_impl.hpp:211-217, 244-248) */
int oracle_solver_set_surface_map(oracle_solver* s, int64_t n, const int* surf_index_of_level_index);
int oracle_solver_init(oracle_solver* s, const oracle_mat* A);
/* x: in = start iterate, out = solution; b is not modified.  Returns 0 on success
 * (converged), 1 if the convergence check failed, <0 on error. */
int oracle_solver_apply(oracle_solver* s, double* x, const double* b, int vblock);
int oracle_solver_steps(const oracle_solver* s);
/* defect history: entry 0 = initial defect; returns number of entries copied */
int oracle_solver_history(const oracle_solver* s, double* out, int cap);
/* one application of the configured preconditioner alone: c = M^{-1} d */
int oracle_precond_apply(oracle_solver* s, double* c, const double* d, int vblock);

#ifdef __cplusplus
}
#endif
#endif
