/*
 * synth.cpp — structured-grid FV1/Q1 problem hierarchies (see synth.h).
 * Host-only input generation; OpenMP over rows.
 */
#include "synth.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct Crs {
	int64_t nrows = 0, ncols = 0;
	int block = 1;
	std::vector<int64_t> rowptr;
	std::vector<int> cols;
	std::vector<double> vals;
	void view(synth_crs* o) const {
		o->nrows = nrows; o->ncols = ncols; o->nnz = (int64_t)cols.size();
		o->block = block; o->rowptr = rowptr.data(); o->cols = cols.data(); o->vals = vals.data();
	}
};

struct Level {
	int lev = 0;
	int n[3] = {1, 1, 1};      // local nodes per direction
	int ne[3] = {0, 0, 0};     // local elements per direction
	int64_t e0[3] = {0, 0, 0}; // global element offset of this rank's box
	int64_t NE[3] = {0, 0, 0}; // global elements per direction
	double h = 1.0;            // cell size
	int64_t nn = 0;            // local node count
	std::vector<int64_t> ord;    // lex -> dof
	std::vector<int64_t> lexof;  // dof -> lex
	std::vector<unsigned char> dir; // per dof
	std::vector<int64_t> gid;    // per dof: global lexicographic node id
	Crs A, P, R;                 // P,R: to/from level lev-1 (empty on lowest level)
};

} // namespace

struct synth_problem {
	synth_desc d;
	int block = 1;
	int nc = 8; // corners per element
	std::vector<std::unique_ptr<Level>> lv; // index = lev - base_lev
	std::vector<double> rhs, exact;
	Level& L(int lev) { return *lv[lev - d.base_lev]; }
	const Level& L(int lev) const { return *lv[lev - d.base_lev]; }
};

namespace {

inline int64_t lexidx(const Level& L, int i, int j, int k) {
	return (int64_t)i + (int64_t)L.n[0] * ((int64_t)j + (int64_t)L.n[1] * k);
}

// ---- element matrices -------------------------------------------------------

// FV1 convection-diffusion on a box cell with edge lengths hd[]; corners numbered
// a = ax + 2*ay (+ 4*az).  Follows the SCVF construction of fv1_geom.cpp:118-139:
// one sub-control-volume face per element edge, integration point = average of the
// SCVF corners, normal scaled by the SCVF area and oriented from -> to.
// Diffusive flux -eps * grad(N_j)(ip).n is added to row `from`, subtracted from row
// `to`; convective flux uses full upwinding.
void fv1_element(int dim, const double hd[3], double eps, const double vel[3], std::vector<double>& Ke)
{
	const int nc = 1 << dim;
	Ke.assign((size_t)nc * nc, 0.0);
	for (int d = 0; d < dim; ++d) {
		const int nOther = dim - 1;
		int od[2] = {0, 0};
		for (int q = 0, c = 0; q < dim; ++q) if (q != d) od[c++] = q;
		for (int st = 0; st < (1 << nOther); ++st) {
			int sv[2] = {st & 1, (st >> 1) & 1};
			// from / to corners
			int a = 0;
			for (int q = 0; q < nOther; ++q) a |= sv[q] << od[q];
			const int b = a | (1 << d);
			double area = 1.0;
			double ipo[2] = {0, 0};
			for (int q = 0; q < nOther; ++q) {
				area *= 0.5 * hd[od[q]];
				ipo[q] = sv[q] ? 0.75 : 0.25;
			}
			for (int j = 0; j < nc; ++j) {
				double g = ((j >> d) & 1) ? 1.0 / hd[d] : -1.0 / hd[d];
				for (int q = 0; q < nOther; ++q)
					g *= ((j >> od[q]) & 1) ? ipo[q] : (1.0 - ipo[q]);
				const double f = -eps * g * area;
				Ke[(size_t)a * nc + j] += f;
				Ke[(size_t)b * nc + j] -= f;
			}
			const double bn = vel[d] * area;
			if (bn != 0.0) {
				const double up = bn > 0 ? bn : 0.0, dn = bn < 0 ? bn : 0.0;
				Ke[(size_t)a * nc + a] += up;
				Ke[(size_t)a * nc + b] += dn;
				Ke[(size_t)b * nc + a] -= up;
				Ke[(size_t)b * nc + b] -= dn;
			}
		}
	}
}

// Q1 (trilinear / bilinear) linear elasticity, 2^dim Gauss points, block = dim.
// Ke[(a*B+i)*(nc*B) + (b*B+j)]
void q1_elasticity_element(int dim, const double hd[3], double E, double nu, std::vector<double>& Ke)
{
	const int nc = 1 << dim, B = dim, N = nc * B;
	Ke.assign((size_t)N * N, 0.0);
	const double lam = E * nu / ((1 + nu) * (1 - 2 * nu)), mu = E / (2 * (1 + nu));
	const double gp[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
	double vol = 1.0;
	for (int d = 0; d < dim; ++d) vol *= hd[d];
	const double w = vol / nc;
	for (int q = 0; q < nc; ++q) {
		double xi[3] = {gp[q & 1], gp[(q >> 1) & 1], gp[(q >> 2) & 1]};
		double G[8][3];
		for (int a = 0; a < nc; ++a)
			for (int d = 0; d < dim; ++d) {
				double g = ((a >> d) & 1) ? 1.0 / hd[d] : -1.0 / hd[d];
				for (int o = 0; o < dim; ++o)
					if (o != d) g *= ((a >> o) & 1) ? xi[o] : (1.0 - xi[o]);
				G[a][d] = g;
			}
		for (int a = 0; a < nc; ++a)
			for (int b = 0; b < nc; ++b) {
				double gg = 0;
				for (int d = 0; d < dim; ++d) gg += G[a][d] * G[b][d];
				for (int i = 0; i < B; ++i)
					for (int j = 0; j < B; ++j) {
						double v = lam * G[a][i] * G[b][j] + mu * G[a][j] * G[b][i];
						if (i == j) v += mu * gg;
						Ke[(size_t)(a * B + i) * N + (b * B + j)] += w * v;
					}
			}
	}
}

// ---- level geometry -----------------------------------------------------------

void build_order(const synth_problem& P, Level& L, const Level* coarser)
{
	L.ord.resize(L.nn);
	L.lexof.resize(L.nn);
	if (P.d.order == SYNTH_ORDER_LEX || !coarser) {
		for (int64_t i = 0; i < L.nn; ++i) { L.ord[i] = i; L.lexof[i] = i; }
		return;
	}
	// UG4-like hierarchical numbering (global_multi_grid_refiner.cpp:249-448):
	// copies of the coarser level's vertices first (in their order), then edge-,
	// face- and volume-midpoints.
	int64_t next = coarser->nn;
	for (int k = 0; k < L.n[2]; k += 2)
		for (int j = 0; j < L.n[1]; j += 2)
			for (int i = 0; i < L.n[0]; i += 2)
				L.ord[lexidx(L, i, j, k)] = coarser->ord[lexidx(*coarser, i / 2, j / 2, k / 2)];
	for (int type = 1; type <= 3; ++type)
		for (int k = 0; k < L.n[2]; ++k)
			for (int j = 0; j < L.n[1]; ++j)
				for (int i = 0; i < L.n[0]; ++i)
					if ((i & 1) + (j & 1) + (k & 1) == type) L.ord[lexidx(L, i, j, k)] = next++;
	for (int64_t i = 0; i < L.nn; ++i) L.lexof[L.ord[i]] = i;
}

inline bool is_dirichlet(const synth_problem& P, const Level& L, int i, int j, int k)
{
	const int ii[3] = {i, j, k};
	if (P.d.problem == SYNTH_ELASTICITY) return L.e0[0] + i == 0;
	for (int d = 0; d < P.d.dim; ++d) {
		const int64_t g = L.e0[d] + ii[d];
		if (g == 0 || g == L.NE[d]) return true;
	}
	return false;
}

void build_geometry(const synth_problem& P, Level& L, const Level* coarser)
{
	const synth_desc& d = P.d;
	L.h = std::ldexp(1.0, -L.lev);
	L.nn = 1;
	for (int q = 0; q < 3; ++q) {
		if (q < d.dim) {
			L.NE[q] = (int64_t)d.base[q] << L.lev;
			L.ne[q] = (int)(L.NE[q] / d.part[q]);
			L.e0[q] = (int64_t)d.coord[q] * L.ne[q];
			L.n[q] = L.ne[q] + 1;
		} else { L.NE[q] = 0; L.ne[q] = 0; L.e0[q] = 0; L.n[q] = 1; }
		L.nn *= L.n[q];
	}
	build_order(P, L, coarser);
	L.dir.resize(L.nn);
	L.gid.resize(L.nn);
	const int64_t GN0 = L.NE[0] + 1, GN1 = L.NE[1] + 1;
#pragma omp parallel for collapse(2) schedule(static)
	for (int k = 0; k < L.n[2]; ++k)
		for (int j = 0; j < L.n[1]; ++j)
			for (int i = 0; i < L.n[0]; ++i) {
				const int64_t dof = L.ord[lexidx(L, i, j, k)];
				L.dir[dof] = is_dirichlet(P, L, i, j, k);
				L.gid[dof] = (L.e0[0] + i) + GN0 * ((L.e0[1] + j) + GN1 * (L.e0[2] + k));
			}
}

// ---- matrix assembly ----------------------------------------------------------

void assemble_matrix(const synth_problem& P, Level& L)
{
	const synth_desc& d = P.d;
	const int dim = d.dim, nc = 1 << dim, B = P.block, BB = B * B;
	double hd[3] = {L.h, L.h, L.h};
	std::vector<double> Ke;
	if (d.problem == SYNTH_ELASTICITY) q1_elasticity_element(dim, hd, d.E, d.nu, Ke);
	else {
		double vel[3] = {0, 0, 0};
		if (d.problem == SYNTH_CONVDIFF) { vel[0] = d.vel[0]; vel[1] = d.vel[1]; vel[2] = d.vel[2]; }
		fv1_element(dim, hd, d.problem == SYNTH_POISSON ? 1.0 : d.eps, vel, Ke);
	}
	const int KN = nc * B;
	Crs& A = L.A;
	A.nrows = A.ncols = L.nn; A.block = B;
	A.rowptr.assign(L.nn + 1, 0);
#pragma omp parallel for schedule(static)
	for (int64_t dof = 0; dof < L.nn; ++dof) {
		int64_t lex = L.lexof[dof];
		const int i = (int)(lex % L.n[0]); lex /= L.n[0];
		const int j = (int)(lex % L.n[1]); const int k = (int)(lex / L.n[1]);
		const int ii[3] = {i, j, k};
		int64_t cnt = 1;
		for (int q = 0; q < dim; ++q) cnt *= 1 + (ii[q] > 0) + (ii[q] < L.n[q] - 1);
		A.rowptr[dof + 1] = cnt;
	}
	for (int64_t r = 0; r < L.nn; ++r) A.rowptr[r + 1] += A.rowptr[r];
	const int64_t nnz = A.rowptr[L.nn];
	A.cols.resize(nnz);
	A.vals.resize((size_t)nnz * BB);
#pragma omp parallel for schedule(static)
	for (int64_t dof = 0; dof < L.nn; ++dof) {
		int64_t lex = L.lexof[dof];
		const int i = (int)(lex % L.n[0]); lex /= L.n[0];
		const int j = (int)(lex % L.n[1]); const int k = (int)(lex / L.n[1]);
		const int ii[3] = {i, j, k};
		double acc[27 * 9];
		bool present[27];
		std::fill(acc, acc + 27 * BB, 0.0);
		std::fill(present, present + 27, false);
		// adjacent elements in lexicographic order; only those of this rank's box
		// contribute (additive matrix, parallel_matrix_impl.h:88-93)
		for (int eo = 0; eo < nc; ++eo) {
			int e[3] = {0, 0, 0}, aLoc = 0;
			bool ok = true;
			for (int q = 0; q < dim; ++q) {
				const int lower = (eo >> q) & 1; // 0: element at ii-1 (node is its upper corner)
				e[q] = ii[q] - 1 + lower;
				if (e[q] < 0 || e[q] >= L.ne[q]) { ok = false; break; }
				if (!lower) aLoc |= 1 << q;
			}
			if (!ok) continue;
			for (int b = 0; b < nc; ++b) {
				int slot = 0, mul = 1;
				for (int q = 0; q < 3; ++q) {
					int off = 0;
					if (q < dim) off = (e[q] + ((b >> q) & 1)) - ii[q];
					slot += (off + 1) * mul; mul *= 3;
				}
				present[slot] = true;
				for (int c = 0; c < B; ++c)
					for (int r = 0; r < B; ++r)
						acc[slot * BB + r + B * c] += Ke[(size_t)(aLoc * B + r) * KN + (b * B + c)];
			}
		}
		const bool dir = L.dir[dof];
		struct Ent { int col; int slot; };
		Ent ent[27]; int ne = 0;
		for (int slot = 0; slot < 27; ++slot) {
			if (!present[slot]) continue;
			const int di = slot % 3 - 1, dj = (slot / 3) % 3 - 1, dk = slot / 9 - 1;
			ent[ne++] = Ent{(int)L.ord[lexidx(L, i + di, j + dj, k + dk)], slot};
		}
		std::sort(ent, ent + ne, [](const Ent& a, const Ent& b) { return a.col < b.col; });
		int64_t p = A.rowptr[dof];
		for (int q = 0; q < ne; ++q, ++p) {
			A.cols[p] = ent[q].col;
			double* v = &A.vals[(size_t)p * BB];
			if (!dir) std::memcpy(v, acc + ent[q].slot * BB, sizeof(double) * BB);
			else {
				// SetDirichletRow: zero row, unit diagonal, pattern retained
				for (int t = 0; t < BB; ++t) v[t] = 0.0;
				if (ent[q].slot == 13) for (int t = 0; t < B; ++t) v[t + B * t] = 1.0;
			}
		}
	}
}

// ---- transfers ----------------------------------------------------------------

void assemble_transfer(const synth_problem& Pb, Level& F, const Level& C)
{
	const int dim = Pb.d.dim;
	Crs& P = F.P;
	P.nrows = F.nn; P.ncols = C.nn; P.block = 1;
	P.rowptr.assign(F.nn + 1, 0);
#pragma omp parallel for schedule(static)
	for (int64_t dof = 0; dof < F.nn; ++dof) {
		int64_t lex = F.lexof[dof];
		const int ii[3] = {(int)(lex % F.n[0]), (int)((lex / F.n[0]) % F.n[1]), (int)(lex / ((int64_t)F.n[0] * F.n[1]))};
		int cnt = 1;
		for (int q = 0; q < dim; ++q) cnt *= 1 + (ii[q] & 1);
		P.rowptr[dof + 1] = cnt;
	}
	for (int64_t r = 0; r < F.nn; ++r) P.rowptr[r + 1] += P.rowptr[r];
	P.cols.resize(P.rowptr[F.nn]);
	P.vals.resize(P.rowptr[F.nn]);
#pragma omp parallel for schedule(static)
	for (int64_t dof = 0; dof < F.nn; ++dof) {
		int64_t lex = F.lexof[dof];
		const int ii[3] = {(int)(lex % F.n[0]), (int)((lex / F.n[0]) % F.n[1]), (int)(lex / ((int64_t)F.n[0] * F.n[1]))};
		int nodd = 0;
		for (int q = 0; q < dim; ++q) nodd += ii[q] & 1;
		// vertex parent 1, edge 1/2, face 1/4, volume 1/8 (std_transfer_impl.h:128-160)
		double w = std::ldexp(1.0, -nodd);
		// Dirichlet fine rows are zeroed, vertex parents re-injected
		// (lagrange_dirichlet_boundary_impl.h:596-611)
		if (F.dir[dof]) w = nodd == 0 ? 1.0 : 0.0;
		struct Ent { int col; };
		int cols[8]; int ne = 0;
		for (int c = 0; c < (1 << nodd); ++c) {
			int cc[3] = {0, 0, 0}, bit = 0;
			for (int q = 0; q < 3; ++q) {
				if (q < dim && (ii[q] & 1)) { cc[q] = (ii[q] - 1) / 2 + ((c >> bit) & 1); ++bit; }
				else cc[q] = ii[q] / 2;
			}
			cols[ne++] = (int)C.ord[lexidx(C, cc[0], cc[1], cc[2])];
		}
		std::sort(cols, cols + ne);
		int64_t p = P.rowptr[dof];
		for (int q = 0; q < ne; ++q, ++p) { P.cols[p] = cols[q]; P.vals[p] = w; }
	}
	// R = P^T (explicit zeros kept, rows sorted), std_transfer_impl.h:694-695
	Crs& R = F.R;
	R.nrows = C.nn; R.ncols = F.nn; R.block = 1;
	R.rowptr.assign(C.nn + 1, 0);
	const int64_t nnz = (int64_t)P.cols.size();
	for (int64_t p = 0; p < nnz; ++p) R.rowptr[P.cols[p] + 1]++;
	for (int64_t r = 0; r < C.nn; ++r) R.rowptr[r + 1] += R.rowptr[r];
	R.cols.resize(nnz); R.vals.resize(nnz);
	std::vector<int64_t> fill(R.rowptr.begin(), R.rowptr.end() - 1);
	for (int64_t f = 0; f < F.nn; ++f)
		for (int64_t p = P.rowptr[f]; p < P.rowptr[f + 1]; ++p) {
			const int64_t q = fill[P.cols[p]]++;
			R.cols[q] = (int)f; R.vals[q] = P.vals[p];
		}
	// coarse Dirichlet rows: zero, then inject the coinciding fine DoF
	// (lagrange_dirichlet_boundary_impl.h:677-745)
#pragma omp parallel for schedule(static)
	for (int64_t c = 0; c < C.nn; ++c) {
		if (!C.dir[c]) continue;
		int64_t lex = C.lexof[c];
		const int ci[3] = {(int)(lex % C.n[0]), (int)((lex / C.n[0]) % C.n[1]), (int)(lex / ((int64_t)C.n[0] * C.n[1]))};
		const int fcol = (int)F.ord[lexidx(F, 2 * ci[0], Pb.d.dim > 1 ? 2 * ci[1] : 0, Pb.d.dim > 2 ? 2 * ci[2] : 0)];
		for (int64_t p = R.rowptr[c]; p < R.rowptr[c + 1]; ++p)
			R.vals[p] = R.cols[p] == fcol ? 1.0 : 0.0;
	}
}

void build_rhs(synth_problem& Pb)
{
	const synth_desc& d = Pb.d;
	const Level& L = Pb.L(d.num_refs);
	const int dim = d.dim, B = Pb.block;
	Pb.rhs.assign((size_t)L.nn * B, 0.0);
	Pb.exact.assign((size_t)L.nn * B, 0.0);
	const double pi = 3.14159265358979323846;
	const double scv = std::ldexp(std::pow(L.h, dim), -dim); // per adjacent element
	// Poisson: u = 1/2 (u1 + u2), u1 = prod sin(pi x_q / D_q) the fundamental mode of the whole domain
	// [0, D_0] x .. (D_q = base[q] unit cells), u2 = prod sin(pi x_q) the fundamental mode of one base cell.
	// On the unit cube both coincide (u = u1 = u2 bit for bit: s + s and 0.5 * are exact).  On a domain of several
	// base cells — the partitioned runs, one cell per rank — u2 alone would be antisymmetric about every partition
	// plane, so that solution, defect and corrections vanish on all interfaces and the exchanges only ever add
	// +v and -v; with u1 the interface values are O(1) and the solution is neither symmetric nor antisymmetric.
	double lam1 = 0.0;
	for (int q = 0; q < dim; ++q) lam1 += 1.0 / ((double)d.base[q] * (double)d.base[q]);
#pragma omp parallel for schedule(static)
	for (int64_t dof = 0; dof < L.nn; ++dof) {
		int64_t lex = L.lexof[dof];
		const int ii[3] = {(int)(lex % L.n[0]), (int)((lex / L.n[0]) % L.n[1]), (int)(lex / ((int64_t)L.n[0] * L.n[1]))};
		int nel = 1;
		double s = 1.0, s1 = 1.0;
		for (int q = 0; q < dim; ++q) {
			nel *= (ii[q] > 0) + (ii[q] < L.n[q] - 1);
			const double xq = (double)(L.e0[q] + ii[q]) * L.h;
			s *= std::sin(pi * xq);
			s1 *= std::sin(pi * xq / (double)d.base[q]);
		}
		const double vol = nel * scv;
		if (d.problem == SYNTH_POISSON) {
			Pb.exact[dof] = 0.5 * (s1 + s);
			// FV1 source: f evaluated at the vertex times the SCV volume (fv1_geom.h:314)
			if (!L.dir[dof]) Pb.rhs[dof] = (0.5 * (lam1 * pi * pi * s1 + dim * pi * pi * s)) * vol;
		} else if (d.problem == SYNTH_CONVDIFF) {
			if (!L.dir[dof]) Pb.rhs[dof] = vol;
		} else {
			if (!L.dir[dof]) Pb.rhs[(size_t)dof * B + (B - 1)] = -vol;
		}
	}
}

} // namespace

extern "C" {

const char* synth_last_error(void) { return g_err.c_str(); }

int synth_create(const synth_desc* d, synth_problem** out)
{
	*out = nullptr;
	if (d->dim != 2 && d->dim != 3) { g_err = "synth: dim must be 2 or 3"; return 1; }
	if (d->base_lev < 0 || d->base_lev > d->num_refs) { g_err = "synth: need 0 <= base_lev <= num_refs"; return 1; }
	if (d->problem < 0 || d->problem > 2) { g_err = "synth: unknown problem"; return 1; }
	for (int q = 0; q < d->dim; ++q) {
		if (d->base[q] < 1 || d->part[q] < 1 || d->coord[q] < 0 || d->coord[q] >= d->part[q]) {
			g_err = "synth: bad base/part/coord"; return 1;
		}
		if ((((int64_t)d->base[q]) << d->base_lev) % d->part[q]) {
			g_err = "synth: process grid must divide the base-level element counts"; return 1;
		}
	}
	std::unique_ptr<synth_problem> P(new synth_problem);
	P->d = *d;
	for (int q = d->dim; q < 3; ++q) { P->d.base[q] = 0; P->d.part[q] = 1; P->d.coord[q] = 0; }
	P->block = d->problem == SYNTH_ELASTICITY ? d->dim : 1;
	P->nc = 1 << d->dim;
	for (int lev = d->base_lev; lev <= d->num_refs; ++lev) {
		P->lv.emplace_back(new Level);
		Level& L = *P->lv.back();
		L.lev = lev;
		const Level* coarser = lev > d->base_lev ? &P->L(lev - 1) : nullptr;
		build_geometry(*P, L, coarser);
		if (L.nn > 2000000000LL) { g_err = "synth: level too large for int32 columns"; return 1; }
		assemble_matrix(*P, L);
		if (coarser) assemble_transfer(*P, L, *coarser);
	}
	build_rhs(*P);
	*out = P.release();
	return 0;
}

void synth_destroy(synth_problem* p) { delete p; }

int synth_block(const synth_problem* p) { return p->block; }

static bool lev_ok(const synth_problem* p, int lev)
{
	if (lev < p->d.base_lev || lev > p->d.num_refs) { g_err = "synth: level out of range"; return false; }
	return true;
}

int synth_level_dims(const synth_problem* p, int lev, int dims[3])
{
	if (!lev_ok(p, lev)) return 1;
	for (int q = 0; q < 3; ++q) dims[q] = p->L(lev).n[q];
	return 0;
}
int synth_level_matrix(const synth_problem* p, int lev, synth_crs* out)
{
	if (!lev_ok(p, lev)) return 1;
	p->L(lev).A.view(out); return 0;
}
int synth_prolongation(const synth_problem* p, int lev, synth_crs* out)
{
	if (!lev_ok(p, lev) || lev == p->d.base_lev) { g_err = "synth: no transfer below base level"; return 1; }
	p->L(lev).P.view(out); return 0;
}
int synth_restriction(const synth_problem* p, int lev, synth_crs* out)
{
	if (!lev_ok(p, lev) || lev == p->d.base_lev) { g_err = "synth: no transfer below base level"; return 1; }
	p->L(lev).R.view(out); return 0;
}
int synth_rhs(const synth_problem* p, const double** b, int64_t* n)
{ *b = p->rhs.data(); *n = (int64_t)p->rhs.size(); return 0; }
int synth_exact(const synth_problem* p, const double** u, int64_t* n)
{ *u = p->exact.data(); *n = (int64_t)p->exact.size(); return 0; }
int synth_dirichlet(const synth_problem* p, int lev, const unsigned char** f, int64_t* n)
{
	if (!lev_ok(p, lev)) return 1;
	*f = p->L(lev).dir.data(); *n = p->L(lev).nn; return 0;
}
int synth_global_ids(const synth_problem* p, int lev, const int64_t** g, int64_t* n)
{
	if (!lev_ok(p, lev)) return 1;
	*g = p->L(lev).gid.data(); *n = p->L(lev).nn; return 0;
}
int synth_dof_to_lex(const synth_problem* p, int lev, const int64_t** m, int64_t* n)
{
	if (!lev_ok(p, lev)) return 1;
	*m = p->L(lev).lexof.data(); *n = p->L(lev).nn; return 0;
}

} // extern "C"
