/*
 * synth.h — synthetic structured-grid problem hierarchies (host, C ABI).
 *
 * Produces what ugcore's DomainDiscretization + StdTransfer hand to the solve
 * path after assembly (SURVEY.md §8d "Synthetic inputs"): per level a CRS
 * matrix (sorted columns, explicit zeros kept in Dirichlet rows), the P1
 * prolongation P, the restriction R = P^T with Dirichlet adjustment, the
 * right-hand side.  The SAME arrays feed the CPU oracle and the GPU path.
 *
 * This is input generation, not the product hot path and not the oracle.
 *
 * Reference behaviour followed (file:line under /root/reference):
 *   FV1 geometry (SCVF = [edge-mid, face-mids, centre], ip = corner average)
 *       ugbase/lib_disc/spatial_disc/disc_util/fv1_geom.cpp:118-139,276-289
 *   Dirichlet rows: identity, pattern retained
 *       ugbase/lib_algebra/algebra_common/sparsematrix_util.h:850-861
 *   P1 prolongation weights 1, 1/2, 1/4, 1/8
 *       ugbase/lib_disc/operator/linear_operator/std_transfer_impl.h:128-160
 *   Dirichlet adjustment of P and R
 *       ugbase/lib_disc/spatial_disc/constraints/dirichlet_boundary/
 *       lagrange_dirichlet_boundary_impl.h:543-611, 677-745
 *   R = set_as_transpose_of(P) (explicit zeros kept)
 *       ugbase/lib_disc/operator/linear_operator/std_transfer_impl.h:694-695
 *   parallel model: element-wise partition, additive matrices
 *       ugbase/lib_algebra/parallelization/parallel_matrix_impl.h:88-114
 */
#ifndef UG4B200_SYNTH_H
#define UG4B200_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { SYNTH_POISSON = 0, SYNTH_CONVDIFF = 1, SYNTH_ELASTICITY = 2 };
enum { SYNTH_ORDER_LEX = 0, SYNTH_ORDER_HIER = 1 };

typedef struct synth_desc {
	int dim;            /* 2 or 3 */
	int base[3];        /* global base-grid elements per direction (level 0) */
	int num_refs;       /* top level = num_refs; level l has base*2^l elements/dir */
	int base_lev;       /* lowest level generated (>= 0) */
	int problem;        /* SYNTH_* */
	int order;          /* SYNTH_ORDER_* (per-rank local DoF numbering) */
	double eps;         /* diffusion coefficient (POISSON: 1) */
	double vel[3];      /* convection velocity (CONVDIFF) */
	double E, nu;       /* elasticity */
	int part[3];        /* process grid (1,1,1 = serial); must divide base*2^base_lev */
	int coord[3];       /* this rank's position in the process grid */
} synth_desc;

typedef struct synth_problem synth_problem;

/* CRS view; pointers stay valid until synth_destroy. vals holds block*block
 * doubles per entry, column-major inside a block (FixedArray2 default,
 * ugbase/lib_algebra/small_algebra/storage/fixed_array_impl.h:182-203). */
typedef struct synth_crs {
	int64_t nrows, ncols, nnz;
	int block;
	const int64_t* rowptr;
	const int* cols;
	const double* vals;
} synth_crs;

int  synth_create(const synth_desc* d, synth_problem** out);
void synth_destroy(synth_problem* p);
const char* synth_last_error(void);

int synth_block(const synth_problem* p);
/* number of local nodes per direction on level lev */
int synth_level_dims(const synth_problem* p, int lev, int dims[3]);
int synth_level_matrix(const synth_problem* p, int lev, synth_crs* out);
/* P: level lev-1 -> lev (rows = fine nodes, scalar entries also for block problems) */
int synth_prolongation(const synth_problem* p, int lev, synth_crs* out);
/* R: level lev -> lev-1 */
int synth_restriction(const synth_problem* p, int lev, synth_crs* out);
/* top-level right-hand side (block*n doubles, additive in parallel) */
int synth_rhs(const synth_problem* p, const double** b, int64_t* n);
/* analytic solution sampled at the top-level nodes (Poisson only, else zeros) */
int synth_exact(const synth_problem* p, const double** u, int64_t* n);
/* per node: 1 if Dirichlet */
int synth_dirichlet(const synth_problem* p, int lev, const unsigned char** f, int64_t* n);
/* global (serial, lexicographic) node id of every local DoF of level lev */
int synth_global_ids(const synth_problem* p, int lev, const int64_t** g, int64_t* n);
/* local DoF index -> local lexicographic node index (identity for ORDER_LEX) */
int synth_dof_to_lex(const synth_problem* p, int lev, const int64_t** m, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif
