// solvers.h — CG, BiCGStab, LinearSolver and the base solvers for the GPU algebra.
//
//   CG            ugbase/lib_algebra/operator/linear_solver/cg.h:103-242
//   BiCGStab      ugbase/lib_algebra/operator/linear_solver/bicgstab.h:112-383
//   LinearSolver  ugbase/lib_algebra/operator/linear_solver/linear_solver.h:114-196
//   GMRES         ugbase/lib_algebra/operator/linear_solver/gmres.h:64-351
//   LU            ugbase/lib_algebra/operator/linear_solver/lu.h:122-380
//                 (dense kernels no_lapack/lu_decomp.h:45-75, 160-195)
//
// CG keeps every scalar (rho, lambda, alpha, beta, the StdConvCheck state and the defect
// history) on the device: an iteration is a fixed sequence of stream-ordered launches
// with no host round trip, captured once into a CUDA graph and replayed.  The host polls
// the device convergence flag one iteration late; iterations queued past convergence are
// no-ops (ug4b200_set_guard), so x, the iteration count and the history are exactly those
// of the reference loop.
#pragma once
#include "preconditioners.h"
#include <cmath>

namespace ug {

namespace detail {
/// device scalars + convergence state of one Krylov solve
struct KrylovDeviceState {
	enum { RHO_OLD = 0, LAMBDA, ALPHA, RHO, BETA, TMP, OMEGA, NUM };
	double* s = nullptr;             // NUM doubles
	ug4b200_conv_state* conv = nullptr;
	double* history = nullptr;
	int historyCap = 0;
	ug4b200_conv_state* pinned = nullptr; // 2 slots
	void* ev[2] = {nullptr, nullptr};
	void ensure(int cap)
	{
		ug4b200_ctx* c = GPUManager::ctx();
		if (!s) {
			s = (double*)GPUManager::alloc_bytes(sizeof(double) * NUM);
			conv = (ug4b200_conv_state*)GPUManager::alloc_bytes(sizeof(ug4b200_conv_state));
			UG_GPU_CHECK(ug4b200_host_alloc(c, 2 * sizeof(ug4b200_conv_state), (void**)&pinned));
			UG_GPU_CHECK(ug4b200_event_create(c, &ev[0]));
			UG_GPU_CHECK(ug4b200_event_create(c, &ev[1]));
		}
		if (cap > historyCap) {
			GPUManager::free_bytes(history);
			history = (double*)GPUManager::alloc_bytes(sizeof(double) * cap);
			historyCap = cap;
		}
	}
	~KrylovDeviceState()
	{
		ug4b200_ctx* c = GPUManager::ctx_or_null();
		if (!c) return;
		GPUManager::free_bytes(s); GPUManager::free_bytes(conv); GPUManager::free_bytes(history);
		if (pinned) ug4b200_host_free(c, pinned);
		if (ev[0]) ug4b200_event_destroy(c, ev[0]);
		if (ev[1]) ug4b200_event_destroy(c, ev[1]);
	}
};
} // namespace detail

template <typename TVector>
class CG : public IPreconditionedLinearOperatorInverse<TVector> {
  public:
	typedef TVector vector_type;
	typedef IPreconditionedLinearOperatorInverse<TVector> base_type;
	using base_type::convergence_check;
	using base_type::linear_operator;
	using base_type::preconditioner;
	enum { B = TVector::blockSize };

	CG() {}
	explicit CG(SmartPtr<ILinearIterator<vector_type, vector_type> > spPrecond) : base_type(spPrecond) {}
	~CG() { drop_graph(); }
	virtual const char* name() const { return "CG"; }
	/// false: reference-shaped loop with host scalars (one host sync per dot / norm)
	void set_device_resident(bool b) { m_deviceResident = b; }
	void set_use_graph(bool b) { m_useGraph = b; }

	virtual bool apply_return_defect(vector_type& x, vector_type& b)
	{
		UG_GPU_ZONE(CG_apply_return_defect);                          // cg.h:105
		if (x.layouts() && (!b.has_storage_type(PST_ADDITIVE) || !x.has_storage_type(PST_CONSISTENT)))
			UG_THROW("CG::apply_return_defect: Inadequate storage format of Vectors.");
		StdConvCheck<vector_type>* std_cc = dynamic_cast<StdConvCheck<vector_type>*>(convergence_check().get());
		// a debug writer wants x and r after every step (cg.h:124, 195): that is the host-paced loop
		if (m_deviceResident && std_cc && !this->vector_debug_writer_valid()) return apply_device(x, b, *std_cc);
		return apply_host(x, b);
	}

  protected:
	/// debugger output: solution and residual (cg.h:273-280)
	void write_debugXR(vector_type& x, vector_type& r, int loopCnt)
	{
		if (!this->vector_debug_writer_valid()) return;
		char ext[20]; snprintf(ext, 20, "_iter%03d", loopCnt);
		this->write_debug(r, std::string("CG_Residual") + ext + ".vec");
		this->write_debug(x, std::string("CG_Solution") + ext + ".vec");
	}
	// ---- reference-shaped loop, scalars on the host (cg.h:103-242) ----
	bool apply_host(vector_type& x, vector_type& b)
	{
		vector_type& r = b;
		linear_operator()->apply_sub(r, x);
		SmartPtr<vector_type> spQ = r.clone_without_values(); vector_type& q = *spQ;
		SmartPtr<vector_type> spZ = x.clone_without_values(); vector_type& z = *spZ;
		SmartPtr<vector_type> spP = x.clone_without_values(); vector_type& p = *spP;
		write_debugXR(x, r, convergence_check()->step());                  // cg.h:124 (the step count is the check's current one)
		if (preconditioner()) { if (!preconditioner()->apply(z, r)) return false; }
		else z = r;
		if (z.layouts() && !z.change_storage_type(PST_CONSISTENT)) UG_THROW("CG: Cannot convert z to consistent vector.");
		convergence_check()->start(r);
		p = z;
		number rhoOld = z.dotprod(r), rho;
		while (!convergence_check()->iteration_ended()) {
			linear_operator()->apply(q, p);
			number lambda = q.dotprod(p);
			if (lambda == 0.0) { if (p.size()) return false; lambda = 1.0; }
			const number alpha = rhoOld / lambda;
			VecScaleAdd(x, 1.0, x, alpha, p);
			VecScaleAdd(r, 1.0, r, -alpha, q);
			write_debugXR(x, r, convergence_check()->step());              // cg.h:195 (before the check counts the step)
			convergence_check()->update(r);
			if (convergence_check()->iteration_ended()) break;
			if (preconditioner()) { if (!preconditioner()->apply(z, r)) return false; }
			else z = r;
			if (z.layouts() && !z.change_storage_type(PST_CONSISTENT)) UG_THROW("CG: Cannot convert z to consistent vector.");
			rho = z.dotprod(r);
			const number beta = rho / rhoOld;
			VecScaleAdd(p, beta, p, 1.0, z);
			rhoOld = rho;
		}
		return convergence_check()->post();
	}

	// ---- device-resident loop ----
	typedef detail::KrylovDeviceState KS;

	void reduce_fin(const vector_type& a, const vector_type& bvec, ug4b200_fin fin, bool parallel)
	{
		ug4b200_ctx* c = GPUManager::ctx();
		if (!parallel) { UG_GPU_CHECK(ug4b200_vec_dot_ds(c, a.len(), a.dev(), bvec.dev(), fin)); return; }
		// local dot + sum over ranks + finaliser: one kernel over the peer windows, three launches over NCCL
		UG_GPU_CHECK(ug4b200_vec_dot_allreduce_ds(c, a.len(), a.dev(), bvec.dev(), fin, m_ks.s + KS::TMP));
	}
	void norm_fin(vector_type& r, int op, bool parallel)
	{
		ug4b200_fin fin{op, nullptr, nullptr, nullptr, m_ks.conv};
		if (parallel && !r.change_storage_type(PST_UNIQUE)) UG_THROW("CG: cannot make the defect unique");
		reduce_fin(r, r, fin, parallel);
	}
	void precond_apply(vector_type& z, vector_type& r)
	{
		if (preconditioner()) { if (!preconditioner()->apply(z, r)) UG_THROW("CG: Cannot apply preconditioner."); }
		else { UG_GPU_CHECK(ug4b200_vec_copy(GPUManager::ctx(), z.len(), z.dev(), r.dev())); z.set_storage_type(r.get_storage_mask()); }
		if (z.layouts() && !z.change_storage_type(PST_CONSISTENT)) UG_THROW("CG: Cannot convert z to consistent vector.");
	}
	void iteration_body(vector_type& x, vector_type& r, vector_type& q, vector_type& z, vector_type& p, bool parallel)
	{
		ug4b200_ctx* c = GPUManager::ctx();
		double* S = m_ks.s;
		// q = A p ; lambda = (q,p) ; alpha = rhoOld / lambda            (cg.h:166-185)
		ug4b200_fin finL{UG4B200_FIN_A_DIV_R, S + KS::LAMBDA, S + KS::ALPHA, S + KS::RHO_OLD, m_ks.conv};
		typedef MatrixOperator<GPUSparseMatrix<typename matrix_value<B>::type>, vector_type> matop_t;
		matop_t* mop = dynamic_cast<matop_t*>(linear_operator().get());
		if (mop) {
			// parallel: q additive, p consistent -> the local dots add up to (q,p); the kernel's last block
			// sums them over the ranks (parallel_vector_impl.h:366-375)
			UG_GPU_CHECK(ug4b200_matrix_apply_dot_allreduce_ds(c, mop->device(), q.dev(), p.dev(), finL, S + KS::TMP));
			q.set_storage_type(PST_ADDITIVE);
		} else {
			linear_operator()->apply(q, p);
			reduce_fin(q, p, finL, parallel);
		}
		// x += alpha p ; r -= alpha q ; ||r|| ; convergence check          (cg.h:187-199)
		if (!parallel) {
			ug4b200_fin finN{UG4B200_FIN_CONV_UPDATE, nullptr, nullptr, nullptr, m_ks.conv};
			UG_GPU_CHECK(ug4b200_cg_update_ds(c, x.len(), x.dev(), p.dev(), r.dev(), q.dev(), S + KS::ALPHA, finN));
		} else {
			ug4b200_fin junk{UG4B200_FIN_STORE, S + KS::TMP, nullptr, nullptr, nullptr};
			UG_GPU_CHECK(ug4b200_cg_update_ds(c, x.len(), x.dev(), p.dev(), r.dev(), q.dev(), S + KS::ALPHA, junk));
			r.set_storage_type(PST_ADDITIVE);
			norm_fin(r, UG4B200_FIN_CONV_UPDATE, true);
		}
		// z = M^-1 r ; rho = (z,r) ; beta = rho/rhoOld ; p = beta p + z ; rhoOld = rho   (cg.h:206-240)
		precond_apply(z, r);
		ug4b200_fin finR{UG4B200_FIN_R_DIV_A, S + KS::RHO_OLD, S + KS::BETA, S + KS::RHO_OLD, nullptr};
		reduce_fin(z, r, finR, parallel);
		UG_GPU_CHECK(ug4b200_vec_scale_add2_ds(c, p.len(), p.dev(), ug4b200_coef{S + KS::BETA, 1.0}, p.dev(),
		                                       ug4b200_coef{nullptr, 1.0}, z.dev()));
	}

	bool apply_device(vector_type& x, vector_type& b, StdConvCheck<vector_type>& cc)
	{
		ug4b200_ctx* c = GPUManager::ctx();
		const bool parallel = (bool)x.layouts();
		vector_type& r = b;
		linear_operator()->apply_sub(r, x);
		SmartPtr<vector_type> spQ = r.clone_without_values(); vector_type& q = *spQ;
		SmartPtr<vector_type> spZ = x.clone_without_values(); vector_type& z = *spZ;
		SmartPtr<vector_type> spP = x.clone_without_values(); vector_type& p = *spP;
		const int maxSteps = cc.maximum_steps();
		m_ks.ensure(maxSteps + 2);
		UG_GPU_CHECK(ug4b200_conv_init(c, m_ks.conv, maxSteps, cc.minimum_defect(), cc.relative_reduction(), m_ks.history, m_ks.historyCap));
		UG_GPU_CHECK(ug4b200_set_guard(c, nullptr));
		precond_apply(z, r);
		norm_fin(r, UG4B200_FIN_CONV_START, parallel);
		UG_GPU_CHECK(ug4b200_vec_copy(c, p.len(), p.dev(), z.dev()));
		p.set_storage_type(z.get_storage_mask());
		ug4b200_fin finR{UG4B200_FIN_STORE, m_ks.s + KS::RHO_OLD, nullptr, nullptr, nullptr};
		reduce_fin(z, r, finR, parallel);
		UG_GPU_CHECK(ug4b200_set_guard(c, &m_ks.conv->done));

		// graph of one iteration, keyed by the buffers it touches
		const void* key[5] = {x.dev(), r.dev(), q.dev(), z.dev(), p.dev()};
		bool graphOk = m_useGraph;
		if (graphOk && (!m_graph || m_graphGen != GPUManager::generation() || std::memcmp(key, m_graphKey, sizeof(key)) != 0)) {
			drop_graph();
			m_graphGen = GPUManager::generation();
			UG_GPU_CHECK(ug4b200_graph_begin(c));
			try { iteration_body(x, r, q, z, p, parallel); }
			catch (...) { ug4b200_graph* g = nullptr; ug4b200_graph_end(c, &g); ug4b200_graph_destroy(c, g); ug4b200_set_guard(c, nullptr); throw; }
			UG_GPU_CHECK(ug4b200_graph_end(c, &m_graph));
			std::memcpy(m_graphKey, key, sizeof(key));
		}
		int it = 0;
		bool done = false;
		// the start defect may already satisfy the check: poll once before iterating
		UG_GPU_CHECK(ug4b200_d2h_async(c, &m_ks.pinned[0], m_ks.conv, sizeof(ug4b200_conv_state)));
		UG_GPU_CHECK(ug4b200_event_record(c, m_ks.ev[0]));
		for (; it < maxSteps && !done; ++it) {
			if (graphOk) UG_GPU_CHECK(ug4b200_graph_launch(c, m_graph));
			else iteration_body(x, r, q, z, p, parallel);
			const int slot = (it + 1) & 1;
			UG_GPU_CHECK(ug4b200_d2h_async(c, &m_ks.pinned[slot], m_ks.conv, sizeof(ug4b200_conv_state)));
			UG_GPU_CHECK(ug4b200_event_record(c, m_ks.ev[slot]));
			// look at the state of the PREVIOUS point in time (the GPU stays one iteration ahead)
			UG_GPU_CHECK(ug4b200_event_sync(c, m_ks.ev[slot ^ 1]));
			done = m_ks.pinned[slot ^ 1].done != 0;
		}
		UG_GPU_CHECK(ug4b200_set_guard(c, nullptr));
		ug4b200_conv_state fin;
		UG_GPU_CHECK(ug4b200_d2h(c, &fin, m_ks.conv, sizeof(fin)));
		std::vector<number> hist(fin.step + 1);
		UG_GPU_CHECK(ug4b200_d2h(c, hist.data(), m_ks.history, sizeof(double) * hist.size()));
		cc.adopt_device_state(fin, hist);
		r.set_storage_type(PST_ADDITIVE);
		if (fin.status == 4) return false; // lambda == 0 breakdown (cg.h:172-184)
		return cc.post();
	}
	void drop_graph()
	{
		if (m_graph && GPUManager::ctx_or_null()) ug4b200_graph_destroy(GPUManager::ctx_or_null(), m_graph);
		m_graph = nullptr;
	}
	template <int N, int dummy = 0> struct matrix_value { typedef DenseMatrix<FixedArray2<double, N, N> > type; };
	template <int dummy> struct matrix_value<1, dummy> { typedef double type; };

	bool m_deviceResident = true, m_useGraph = true;
	KS m_ks;
	ug4b200_graph* m_graph = nullptr;
	unsigned long long m_graphGen = 0;   // GPUManager::generation() the graph was captured under
	const void* m_graphKey[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

/// BiCGStab, reference-shaped loop (bicgstab.h:112-383); vector work on the device, the
/// handful of scalars per half-step on the host (restart logic needs them there)
template <typename TVector>
class BiCGStab : public IPreconditionedLinearOperatorInverse<TVector> {
  public:
	typedef TVector vector_type;
	typedef IPreconditionedLinearOperatorInverse<TVector> base_type;
	using base_type::convergence_check;
	using base_type::linear_operator;
	using base_type::preconditioner;

	BiCGStab() : m_numRestarts(0), m_minOrtho(0.0) {}
	~BiCGStab() { drop_graph(); if (m_scal) GPUManager::free_bytes(m_scal); }
	virtual const char* name() const { return "BiCGStab"; }
	void set_restart(int numRestarts) { m_numRestarts = numRestarts; }
	void set_min_orthogonality(number minOrtho) { m_minOrtho = minOrtho; }
	/// true (default; measured 33.3 vs 38.6 ms per solve on BASELINE configs[3] at 129^3, identical histories): every
	/// scalar of the iteration (rho, alpha, omega, beta, the convergence state) stays on the device and one iteration
	/// is a CUDA graph, like CG::apply_device; false: reference-shaped host loop (one host sync per dot)
	void set_device_resident(bool b) { m_deviceResident = b; }
	void set_use_graph(bool b) { m_useGraph = b; }

	virtual bool apply_return_defect(vector_type& x, vector_type& b)
	{
		UG_GPU_ZONE(LS_ApplyReturnDefect);                            // bicgstab.h:114
		const bool par = (bool)x.layouts();
		if (par && (!b.has_storage_type(PST_ADDITIVE) || !x.has_storage_type(PST_CONSISTENT)))
			UG_THROW("BiCGStab: Inadequate storage format of Vectors.");
		StdConvCheck<vector_type>* std_cc = dynamic_cast<StdConvCheck<vector_type>*>(convergence_check().get());
		if (m_deviceResident && std_cc && m_minOrtho == 0.0) return apply_device(x, b, *std_cc);
		linear_operator()->apply_sub(b, x);
		vector_type& r = b;
		SmartPtr<vector_type> spR = r.clone_without_values(); vector_type& r0 = *spR;
		SmartPtr<vector_type> spP = r.clone_without_values(); vector_type& p = *spP;
		SmartPtr<vector_type> spV = r.clone_without_values(); vector_type& v = *spV;
		SmartPtr<vector_type> spT = r.clone_without_values(); vector_type& t = *spT;
		SmartPtr<vector_type> spS = r.clone_without_values(); vector_type& s = *spS;
		SmartPtr<vector_type> spQ = x.clone_without_values(); vector_type& q = *spQ;
		convergence_check()->start(r);
		if (par && !r.change_storage_type(PST_UNIQUE)) UG_THROW("BiCGStab: Cannot convert b to unique.");
		number rho = 1, alpha = 1, omega = 1, norm_r0 = 0.0;
		bool bRestart = true;
		while (!convergence_check()->iteration_ended()) {
			if (m_numRestarts > 0 && (convergence_check()->step() % m_numRestarts == 0)) bRestart = true;
			if (bRestart) {
				r0 = r;
				if (par && !r0.change_storage_type(PST_UNIQUE)) UG_THROW("BiCGStab: Cannot convert r to unique vector.");
				p = 0.0; alpha = 0.0; p.set_storage_type(r.get_storage_mask());
				v = 0.0; omega = 1.0; v.set_storage_type(r.get_storage_mask());
				rho = 1.0;
				norm_r0 = convergence_check()->defect();
				bRestart = false;
			}
			const number rhoOld = rho;
			if (!r.size()) rho = 1.0; else rho = r0.dotprod(r);
			const number norm_r = convergence_check()->defect();
			if (std::fabs(rho) / (norm_r * norm_r0) <= m_minOrtho) bRestart = true;
			if (rhoOld == 0.0) return false;
			const number beta = (rho / rhoOld) * (alpha / omega);
			VecScaleAdd(p, 1.0, r, beta, p, -beta * omega, v);
			if (preconditioner()) { if (!preconditioner()->apply(q, p)) return false; }
			else { q = p; if (par && !q.change_storage_type(PST_CONSISTENT)) UG_THROW("BiCGStab: Cannot convert q to consistent vector."); }
			linear_operator()->apply(v, q);
			if (par && !v.change_storage_type(PST_UNIQUE)) UG_THROW("BiCGStab: Cannot convert v to unique vector.");
			if (!v.size()) alpha = 1.0; else alpha = v.dotprod(r0);
			if (alpha == 0.0) return false;
			alpha = rho / alpha;
			VecScaleAdd(x, 1.0, x, alpha, q);
			VecScaleAdd(s, 1.0, r, -alpha, v);
			convergence_check()->update(s);
			if (convergence_check()->iteration_ended()) { r = s; break; }
			if (preconditioner()) { if (!preconditioner()->apply(q, s)) return false; }
			else { q = s; if (par && !q.change_storage_type(PST_CONSISTENT)) UG_THROW("BiCGStab: Cannot convert q to consistent vector."); }
			linear_operator()->apply(t, q);
			if (par && !t.change_storage_type(PST_UNIQUE)) UG_THROW("BiCGStab: Cannot convert t to unique vector.");
			number tt;
			if (!t.size()) tt = 1.0; else tt = t.dotprod(t);
			if (!s.size()) omega = 1.0; else omega = s.dotprod(t);
			if (tt == 0.0) return false;
			omega = omega / tt;
			VecScaleAdd(x, 1.0, x, omega, q);
			VecScaleAdd(r, 1.0, s, -omega, t);
			convergence_check()->update(r);
			if (omega == 0.0) return false;
		}
		return convergence_check()->post();
	}

  protected:
	// ---- device-resident loop: the statements of bicgstab.h:161-380 with device scalars ----
	typedef detail::KrylovDeviceState KS;
	enum { S_RHO = KS::RHO, S_RHO_OLD = KS::RHO_OLD, S_ALPHA = KS::ALPHA, S_OMEGA = KS::OMEGA, S_BETA = KS::BETA, S_TMP = KS::TMP,
	       S_TT = KS::LAMBDA, S_BO = KS::NUM };   // BO = beta * omega lives behind the shared enum (ensure() allocates NUM + 1)

	void reduce_fin(const vector_type& a, const vector_type& bvec, ug4b200_fin fin, bool par)
	{
		ug4b200_ctx* c = GPUManager::ctx();
		if (!par) { UG_GPU_CHECK(ug4b200_vec_dot_ds(c, a.len(), a.dev(), bvec.dev(), fin)); return; }
		UG_GPU_CHECK(ug4b200_vec_dot_allreduce_ds(c, a.len(), a.dev(), bvec.dev(), fin, m_scal + S_TMP));
	}
	void precond_apply(vector_type& q, vector_type& p, bool par)
	{
		if (preconditioner()) { if (!preconditioner()->apply(q, p)) UG_THROW("BiCGStab: Cannot apply preconditioner."); }
		else {
			UG_GPU_CHECK(ug4b200_vec_copy(GPUManager::ctx(), q.len(), q.dev(), p.dev())); q.set_storage_type(p.get_storage_mask());
			if (par && !q.change_storage_type(PST_CONSISTENT)) UG_THROW("BiCGStab: Cannot convert q to consistent vector.");
		}
	}
	/// dest = 1.0 * a + (sign * *coef) * bvec, then ||dest|| -> convergence update   (bicgstab.h:285-291, 362-368)
	void update_and_check(vector_type& dest, const vector_type& a, const double* coef, double sign, const vector_type& bvec, bool par)
	{
		ug4b200_ctx* c = GPUManager::ctx();
		ug4b200_fin fin{UG4B200_FIN_CONV_UPDATE, nullptr, nullptr, nullptr, m_ks.conv};
		const unsigned type = a.get_storage_mask() & bvec.get_storage_mask();
		if (!par) {
			UG_GPU_CHECK(ug4b200_vec_scale_add2_norm_ds(c, dest.len(), dest.dev(), ug4b200_coef{nullptr, 1.0}, a.dev(), ug4b200_coef{coef, sign}, bvec.dev(), fin));
			dest.set_storage_type(type);
			return;
		}
		UG_GPU_CHECK(ug4b200_vec_scale_add2_ds(c, dest.len(), dest.dev(), ug4b200_coef{nullptr, 1.0}, a.dev(), ug4b200_coef{coef, sign}, bvec.dev()));
		dest.set_storage_type(type);
		if (!dest.change_storage_type(PST_UNIQUE)) UG_THROW("BiCGStab: cannot make the defect unique");
		reduce_fin(dest, dest, fin, true);
	}
	void restart_block(vector_type& r, vector_type& r0, vector_type& p, vector_type& v, bool par)
	{
		ug4b200_ctx* c = GPUManager::ctx();
		r0 = r;                                                                     // :167
		if (par && !r0.change_storage_type(PST_UNIQUE)) UG_THROW("BiCGStab: Cannot convert r to unique vector.");
		p.set(0.0); p.set_storage_type(r.get_storage_mask());                       // :176-181
		v.set(0.0); v.set_storage_type(r.get_storage_mask());
		UG_GPU_CHECK(ug4b200_vec_set(c, 1, m_scal + S_ALPHA, 0.0));
		UG_GPU_CHECK(ug4b200_vec_set(c, 1, m_scal + S_OMEGA, 1.0));
		UG_GPU_CHECK(ug4b200_vec_set(c, 1, m_scal + S_RHO, 1.0));                  // :184
	}
	void iteration_body(vector_type& x, vector_type& r, vector_type& r0, vector_type& p, vector_type& v, vector_type& t,
	                    vector_type& s, vector_type& q, bool par)
	{
		ug4b200_ctx* c = GPUManager::ctx();
		double* S = m_scal;
		UG_GPU_CHECK(ug4b200_vec_copy(c, 1, S + S_RHO_OLD, S + S_RHO));             // rhoOld = rho                  :198
		reduce_fin(r0, r, ug4b200_fin{UG4B200_FIN_STORE, S + S_RHO, nullptr, nullptr, nullptr}, par);   // rho = (r0, r) :204
		// beta = (rho / rhoOld) * (alpha / omega); rhoOld == 0 or omega == 0 end the solve (:216-224, :374-381)
		UG_GPU_CHECK(ug4b200_scalar_ratio_conv_ds(c, S + S_BETA, S + S_RHO, S + S_RHO_OLD, S + S_ALPHA, S + S_OMEGA, m_ks.conv));   // :227
		UG_GPU_CHECK(ug4b200_scalar_ratio_ds(c, S + S_BO, S + S_BETA, nullptr, S + S_OMEGA, nullptr));              // beta*omega
		UG_GPU_CHECK(ug4b200_vec_scale_add3_ds(c, p.len(), p.dev(), ug4b200_coef{nullptr, 1.0}, r.dev(), ug4b200_coef{S + S_BETA, 1.0}, p.dev(),
		                                       ug4b200_coef{S + S_BO, -1.0}, v.dev()));                           // :230
		p.set_storage_type(r.get_storage_mask() & p.get_storage_mask() & v.get_storage_mask());
		precond_apply(q, p, par);                                                   // :236-251
		linear_operator()->apply(v, q);                                             // :260
		if (par && !v.change_storage_type(PST_UNIQUE)) UG_THROW("BiCGStab: Cannot convert v to unique vector.");
		reduce_fin(v, r0, ug4b200_fin{UG4B200_FIN_A_DIV_R, S + S_TMP, S + S_ALPHA, S + S_RHO, m_ks.conv}, par);   // alpha = rho / (v, r0)  :272-282
		UG_GPU_CHECK(ug4b200_vec_scale_add2_ds(c, x.len(), x.dev(), ug4b200_coef{nullptr, 1.0}, x.dev(), ug4b200_coef{S + S_ALPHA, 1.0}, q.dev()));   // :285
		update_and_check(s, r, S + S_ALPHA, -1.0, v, par);                          // s = r - alpha v ; check   :288-291
		// (converged here: the guard turns the rest of the iteration into no-ops; r = s is done by the caller, :294-297)
		precond_apply(q, s, par);                                                   // :305-320
		linear_operator()->apply(t, q);                                             // :329
		if (par && !t.change_storage_type(PST_UNIQUE)) UG_THROW("BiCGStab: Cannot convert t to unique vector.");
		reduce_fin(t, t, ug4b200_fin{UG4B200_FIN_STORE, S + S_TT, nullptr, nullptr, nullptr}, par);               // tt = (t, t)   :342
		reduce_fin(s, t, ug4b200_fin{UG4B200_FIN_R_DIV_A, S + S_TMP, S + S_OMEGA, S + S_TT, m_ks.conv}, par);     // omega = (s, t) / tt, tt == 0: breakdown   :344-359
		UG_GPU_CHECK(ug4b200_vec_scale_add2_ds(c, x.len(), x.dev(), ug4b200_coef{nullptr, 1.0}, x.dev(), ug4b200_coef{S + S_OMEGA, 1.0}, q.dev()));   // :362
		update_and_check(r, s, S + S_OMEGA, -1.0, t, par);                          // r = s - omega t ; check   :365-368
	}
	bool apply_device(vector_type& x, vector_type& b, StdConvCheck<vector_type>& cc)
	{
		ug4b200_ctx* c = GPUManager::ctx();
		const bool par = (bool)x.layouts();
		linear_operator()->apply_sub(b, x);
		vector_type& r = b;
		SmartPtr<vector_type> spR = r.clone_without_values(); vector_type& r0 = *spR;
		SmartPtr<vector_type> spP = r.clone_without_values(); vector_type& p = *spP;
		SmartPtr<vector_type> spV = r.clone_without_values(); vector_type& v = *spV;
		SmartPtr<vector_type> spT = r.clone_without_values(); vector_type& t = *spT;
		SmartPtr<vector_type> spS = r.clone_without_values(); vector_type& s = *spS;
		SmartPtr<vector_type> spQ = x.clone_without_values(); vector_type& q = *spQ;
		const int maxSteps = cc.maximum_steps();
		m_ks.ensure(maxSteps + 2);
		if (!m_scal) m_scal = (double*)GPUManager::alloc_bytes(sizeof(double) * (KS::NUM + 1));
		UG_GPU_CHECK(ug4b200_conv_init(c, m_ks.conv, maxSteps, cc.minimum_defect(), cc.relative_reduction(), m_ks.history, m_ks.historyCap));
		UG_GPU_CHECK(ug4b200_set_guard(c, nullptr));
		{   // convergence_check()->start(r), then r unique  (:154-158)
			if (par && !r.change_storage_type(PST_UNIQUE)) UG_THROW("BiCGStab: Cannot convert b to unique.");
			reduce_fin(r, r, ug4b200_fin{UG4B200_FIN_CONV_START, nullptr, nullptr, nullptr, m_ks.conv}, par);
		}
		UG_GPU_CHECK(ug4b200_set_guard(c, &m_ks.conv->done));
		restart_block(r, r0, p, v, par);            // sets the storage types the captured body relies on
		const void* key[8] = {x.dev(), r.dev(), r0.dev(), p.dev(), v.dev(), t.dev(), s.dev(), q.dev()};
		bool graphOk = m_useGraph;
		if (graphOk && (!m_graph || m_graphGen != GPUManager::generation() || std::memcmp(key, m_graphKey, sizeof(key)) != 0)) {
			drop_graph();
			m_graphGen = GPUManager::generation();
			UG_GPU_CHECK(ug4b200_graph_begin(c));
			try { iteration_body(x, r, r0, p, v, t, s, q, par); }
			catch (...) { ug4b200_graph* g = nullptr; ug4b200_graph_end(c, &g); ug4b200_graph_destroy(c, g); ug4b200_set_guard(c, nullptr); throw; }
			UG_GPU_CHECK(ug4b200_graph_end(c, &m_graph));
			std::memcpy(m_graphKey, key, sizeof(key));
		}
		bool done = false;
		UG_GPU_CHECK(ug4b200_d2h_async(c, &m_ks.pinned[0], m_ks.conv, sizeof(ug4b200_conv_state)));
		UG_GPU_CHECK(ug4b200_event_record(c, m_ks.ev[0]));
		const int maxIts = (maxSteps + 1) / 2;
		for (int it = 0; it < maxIts && !done; ++it) {
			// periodic restart (:161-163): an iteration advances the step count by two
			if (it > 0 && m_numRestarts > 0 && (2 * it) % m_numRestarts == 0) restart_block(r, r0, p, v, par);
			if (graphOk) UG_GPU_CHECK(ug4b200_graph_launch(c, m_graph));
			else iteration_body(x, r, r0, p, v, t, s, q, par);
			const int slot = (it + 1) & 1;
			UG_GPU_CHECK(ug4b200_d2h_async(c, &m_ks.pinned[slot], m_ks.conv, sizeof(ug4b200_conv_state)));
			UG_GPU_CHECK(ug4b200_event_record(c, m_ks.ev[slot]));
			UG_GPU_CHECK(ug4b200_event_sync(c, m_ks.ev[slot ^ 1]));
			done = m_ks.pinned[slot ^ 1].done != 0;
		}
		UG_GPU_CHECK(ug4b200_set_guard(c, nullptr));
		ug4b200_conv_state fin;
		UG_GPU_CHECK(ug4b200_d2h(c, &fin, m_ks.conv, sizeof(fin)));
		std::vector<number> hist(fin.step + 1);
		UG_GPU_CHECK(ug4b200_d2h(c, hist.data(), m_ks.history, sizeof(double) * hist.size()));
		cc.adopt_device_state(fin, hist);
		if (fin.step % 2 == 1 && fin.status != 4) r = s;   // the check after the half step ended the iteration (:294-297)
		if (fin.status == 4) return false;          // (v, r0) == 0, rhoOld == 0, tt == 0 or omega == 0: the reference returns false
		return cc.post();
	}
	void drop_graph()
	{
		if (m_graph && GPUManager::ctx_or_null()) ug4b200_graph_destroy(GPUManager::ctx_or_null(), m_graph);
		m_graph = nullptr;
	}

	int m_numRestarts;
	number m_minOrtho;
	bool m_deviceResident = true, m_useGraph = true;
	KS m_ks;
	double* m_scal = nullptr;
	ug4b200_graph* m_graph = nullptr;
	unsigned long long m_graphGen = 0;   // GPUManager::generation() the graph was captured under
	const void* m_graphKey[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

/// GMRES(restart), left preconditioned (gmres.h:104-278), reference-shaped: vector work on the device, the
/// Hessenberg matrix and the Givens rotations (a few dozen scalars) on the host, one sync per dot / norm.
/// As in the reference every cycle runs all `restart` inner steps; with a preconditioner the convergence check
/// sees the true defect once per cycle (:275-276), without one the rotated residual norm of every inner
/// step (:249-251).
template <typename TVector>
class GMRES : public IPreconditionedLinearOperatorInverse<TVector> {
  public:
	typedef TVector vector_type;
	typedef IPreconditionedLinearOperatorInverse<TVector> base_type;
	using base_type::convergence_check;
	using base_type::linear_operator;
	using base_type::preconditioner;

	explicit GMRES(size_t restart) : m_restart(restart) {}
	virtual const char* name() const { return "GMRES"; }

	virtual bool apply_return_defect(vector_type& x, vector_type& b)
	{
		const bool par = (bool)x.layouts();
		if (par && (!b.has_storage_type(PST_ADDITIVE) || !x.has_storage_type(PST_CONSISTENT)))
			UG_THROW("GMRES: Inadequate storage format of Vectors.");
		if (m_restart == 0) UG_THROW("GMRES: restart must be positive");
		SmartPtr<vector_type> spR = b.clone();                       // copy rhs
		linear_operator()->apply_sub(*spR, x);                       // b - A x
		convergence_check()->start(*spR);
		std::vector<SmartPtr<vector_type> > v(m_restart + 1);
		std::vector<std::vector<number> > h(m_restart + 1);
		for (size_t i = 0; i < h.size(); ++i) h[i].resize(m_restart + 1);
		std::vector<number> gamma(m_restart + 1), c(m_restart + 1), s(m_restart + 1);
		while (!convergence_check()->iteration_ended()) {
			if (!v[0]) v[0] = x.clone_without_values();
			// v[0] = M^-1 (b - A x), or reuse the defect vector
			if (preconditioner()) { if (!preconditioner()->apply(*v[0], *spR)) return false; }
			else std::swap(v[0], spR);
			if (par && !v[0]->change_storage_type(PST_UNIQUE)) UG_THROW("GMRES: Cannot convert v0 to consistent vector.");
			gamma[0] = v[0]->norm();
			*v[0] *= 1. / gamma[0];
			size_t numIter = 0;
			for (size_t j = 0; j < m_restart; ++j) {
				numIter = j;
				if (!v[j + 1]) v[j + 1] = x.clone_without_values();
				if (par && !v[j]->change_storage_type(PST_CONSISTENT)) UG_THROW("GMRES: Cannot convert v[" << j + 1 << "] to consistent vector.");
				linear_operator()->apply(*spR, *v[j]);                 // r = A v[j]
				if (preconditioner()) { if (!preconditioner()->apply(*v[j + 1], *spR)) return false; }
				else std::swap(v[j + 1], spR);
				if (par) {
					if (!v[j]->change_storage_type(PST_UNIQUE)) UG_THROW("GMRES: Cannot convert v0 to consistent vector.");
					if (!v[j + 1]->change_storage_type(PST_UNIQUE)) UG_THROW("GMRES: Cannot convert v[" << j << "] to consistent vector.");
				}
				for (size_t i = 0; i <= j; ++i) {
					h[i][j] = v[j + 1]->dotprod(*v[i]);                // h_ij := (v[j+1], v[i])
					VecScaleAppend(*v[j + 1], *v[i], (-1) * h[i][j]);  // v[j+1] -= h_ij v[i]
				}
				h[j + 1][j] = v[j + 1]->norm();
				for (size_t i = 0; i < j; ++i) {                       // apply the previous rotations to the new column
					const number hij = h[i][j], hi1j = h[i + 1][j];
					h[i][j] = c[i + 1] * hij + s[i + 1] * hi1j;
					h[i + 1][j] = s[i + 1] * hij - c[i + 1] * hi1j;
				}
				const number alpha = std::sqrt(h[j][j] * h[j][j] + h[j + 1][j] * h[j + 1][j]);
				s[j + 1] = h[j + 1][j] / alpha;
				c[j + 1] = h[j][j] / alpha;
				h[j][j] = alpha;
				gamma[j + 1] = s[j + 1] * gamma[j];
				gamma[j] = c[j + 1] * gamma[j];
				if (!preconditioner()) convergence_check()->update_defect(gamma[j + 1]);
				*v[j + 1] *= 1. / (h[j + 1][j]);
			}
			for (size_t i = numIter;; --i) {                           // back substitution, x += gamma_i v[i]
				for (size_t j = i + 1; j <= numIter; ++j) gamma[i] -= h[i][j] * gamma[j];
				gamma[i] /= h[i][i];
				VecScaleAppend(x, *v[i], gamma[i]);
				if (i == 0) break;
			}
			*spR = b;                                                  // fresh defect
			linear_operator()->apply_sub(*spR, x);
			if (preconditioner()) convergence_check()->update(*spR);
		}
		return convergence_check()->post();
	}

  protected:
	/// a += s * b with the storage-type reconciliation of gmres.h:326-347
	bool VecScaleAppend(vector_type& a, vector_type& b, number s)
	{
		if (a.layouts()) {
			if (a.has_storage_type(PST_UNIQUE) && b.has_storage_type(PST_UNIQUE)) {}
			else if (a.has_storage_type(PST_CONSISTENT) && b.has_storage_type(PST_CONSISTENT)) {}
			else if (a.has_storage_type(PST_ADDITIVE) && b.has_storage_type(PST_ADDITIVE)) {}
			else { a.change_storage_type(PST_ADDITIVE); b.change_storage_type(PST_ADDITIVE); }
		}
		const unsigned type = a.get_storage_mask();
		VecScaleAdd(a, 1.0, a, s, b);
		a.set_storage_type(type);                                      // the element loop of the reference leaves a's type alone
		return true;
	}
	size_t m_restart;
};

/// LinearSolver (linear_solver.h:114-196): x += B(b - A x) until converged
template <typename TVector>
class LinearSolver : public IPreconditionedLinearOperatorInverse<TVector> {
  public:
	typedef TVector vector_type;
	typedef IPreconditionedLinearOperatorInverse<TVector> base_type;
	using base_type::convergence_check;
	using base_type::linear_operator;
	using base_type::preconditioner;
	virtual const char* name() const { return "Iterative Linear Solver"; }
	~LinearSolver() { drop_graph(); }
	/// true (default, like CG and BiCGStab): the convergence state stays on the device and one iteration (c = B d,
	/// d -= A c, x += c, ||d||) is a CUDA graph; false: the reference-shaped loop, one host sync per step — identical
	/// histories and iterates (tests/test_ilu.py: test_gpu_device_resident_linear_solver_equals_host_loop)
	void set_device_resident(bool b) { m_deviceResident = b; }
	void set_use_graph(bool b) { m_useGraph = b; }
	virtual bool apply_return_defect(vector_type& x, vector_type& b)
	{
		if (x.layouts() && (!b.has_storage_type(PST_ADDITIVE) || !x.has_storage_type(PST_CONSISTENT)))
			UG_THROW("LinearSolver::apply: Inadequate parallel storage format of Vectors.");
		StdConvCheck<vector_type>* std_cc = dynamic_cast<StdConvCheck<vector_type>*>(convergence_check().get());
		if (m_deviceResident && std_cc) return apply_device(x, b, *std_cc);
		vector_type& d = b;
		linear_operator()->apply_sub(d, x);
		SmartPtr<vector_type> spC = x.clone_without_values(); vector_type& c = *spC;
		c.set_storage_type(PST_CONSISTENT);
		convergence_check()->start(d);
		while (!convergence_check()->iteration_ended()) {
			if (preconditioner()) { if (!preconditioner()->apply_update_defect(c, d)) return false; }
			else { c = d; linear_operator()->apply_sub(d, c); }
			x += c;
			convergence_check()->update(d);
		}
		return convergence_check()->post();
	}

  protected:
	typedef detail::KrylovDeviceState KS;
	void norm_fin(vector_type& d, int op, bool par)
	{
		ug4b200_ctx* ctx = GPUManager::ctx();
		ug4b200_fin fin{op, nullptr, nullptr, nullptr, m_ks.conv};
		if (!par) { UG_GPU_CHECK(ug4b200_vec_dot_ds(ctx, d.len(), d.dev(), d.dev(), fin)); return; }
		if (!d.change_storage_type(PST_UNIQUE)) UG_THROW("LinearSolver: cannot make the defect unique");
		UG_GPU_CHECK(ug4b200_vec_dot_allreduce_ds(ctx, d.len(), d.dev(), d.dev(), fin, m_ks.s + KS::TMP));
	}
	void iteration_body(vector_type& x, vector_type& d, vector_type& c, bool par)
	{
		if (preconditioner()) { if (!preconditioner()->apply_update_defect(c, d)) UG_THROW("LinearSolver: Cannot apply preconditioner."); }
		else { c = d; linear_operator()->apply_sub(d, c); }                        // linear_solver.h:160-175
		x += c;                                                                      // :178
		norm_fin(d, UG4B200_FIN_CONV_UPDATE, par);                                   // :181
	}
	bool apply_device(vector_type& x, vector_type& b, StdConvCheck<vector_type>& cc)
	{
		ug4b200_ctx* ctx = GPUManager::ctx();
		const bool par = (bool)x.layouts();
		vector_type& d = b;
		linear_operator()->apply_sub(d, x);
		SmartPtr<vector_type> spC = x.clone_without_values(); vector_type& c = *spC;
		c.set_storage_type(PST_CONSISTENT);
		const int maxSteps = cc.maximum_steps();
		m_ks.ensure(maxSteps + 2);
		UG_GPU_CHECK(ug4b200_conv_init(ctx, m_ks.conv, maxSteps, cc.minimum_defect(), cc.relative_reduction(), m_ks.history, m_ks.historyCap));
		UG_GPU_CHECK(ug4b200_set_guard(ctx, nullptr));
		norm_fin(d, UG4B200_FIN_CONV_START, par);
		UG_GPU_CHECK(ug4b200_set_guard(ctx, &m_ks.conv->done));
		const void* key[3] = {x.dev(), d.dev(), c.dev()};
		const bool graphOk = m_useGraph;
		if (graphOk && (!m_graph || m_graphGen != GPUManager::generation() || std::memcmp(key, m_graphKey, sizeof(key)) != 0)) {
			drop_graph();
			m_graphGen = GPUManager::generation();
			UG_GPU_CHECK(ug4b200_graph_begin(ctx));
			try { iteration_body(x, d, c, par); }
			catch (...) { ug4b200_graph* g = nullptr; ug4b200_graph_end(ctx, &g); ug4b200_graph_destroy(ctx, g); ug4b200_set_guard(ctx, nullptr); throw; }
			UG_GPU_CHECK(ug4b200_graph_end(ctx, &m_graph));
			std::memcpy(m_graphKey, key, sizeof(key));
		}
		bool done = false;
		UG_GPU_CHECK(ug4b200_d2h_async(ctx, &m_ks.pinned[0], m_ks.conv, sizeof(ug4b200_conv_state)));
		UG_GPU_CHECK(ug4b200_event_record(ctx, m_ks.ev[0]));
		for (int it = 0; it < maxSteps && !done; ++it) {
			if (graphOk) UG_GPU_CHECK(ug4b200_graph_launch(ctx, m_graph));
			else iteration_body(x, d, c, par);
			const int slot = (it + 1) & 1;
			UG_GPU_CHECK(ug4b200_d2h_async(ctx, &m_ks.pinned[slot], m_ks.conv, sizeof(ug4b200_conv_state)));
			UG_GPU_CHECK(ug4b200_event_record(ctx, m_ks.ev[slot]));
			UG_GPU_CHECK(ug4b200_event_sync(ctx, m_ks.ev[slot ^ 1]));      // the GPU stays one iteration ahead of the host
			done = m_ks.pinned[slot ^ 1].done != 0;
		}
		UG_GPU_CHECK(ug4b200_set_guard(ctx, nullptr));
		ug4b200_conv_state fin;
		UG_GPU_CHECK(ug4b200_d2h(ctx, &fin, m_ks.conv, sizeof(fin)));
		std::vector<number> hist(fin.step + 1);
		UG_GPU_CHECK(ug4b200_d2h(ctx, hist.data(), m_ks.history, sizeof(double) * hist.size()));
		cc.adopt_device_state(fin, hist);
		d.set_storage_type(PST_ADDITIVE);
		return cc.post();
	}
	void drop_graph()
	{
		if (m_graph && GPUManager::ctx_or_null()) ug4b200_graph_destroy(GPUManager::ctx_or_null(), m_graph);
		m_graph = nullptr;
	}
	bool m_deviceResident = true, m_useGraph = true;
	KS m_ks;
	ug4b200_graph* m_graph = nullptr;
	unsigned long long m_graphGen = 0;   // GPUManager::generation() the graph was captured under
	const void* m_graphKey[3] = {nullptr, nullptr, nullptr};
};

/// LU base solver: dense factorisation on the host at init (assembly side), triangular
/// solves on the device by one CTA (lu.h:122-140, 189-207)
template <typename TAlgebra>
class LU : public ILinearOperatorInverse<typename TAlgebra::vector_type> {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	typedef typename TAlgebra::matrix_type matrix_type;
	typedef MatrixOperator<matrix_type, vector_type> matrix_operator_type;
	enum { B = TAlgebra::blockSize };
	LU() {}
	~LU() { free_dev(); }
	virtual const char* name() const { return "LU"; }
	virtual bool supports_parallel() const { return false; }
	virtual bool init(SmartPtr<ILinearOperator<vector_type> > L)
	{
		m_spOperator = sp_cast_dynamic<matrix_operator_type>(L);
		if (!m_spOperator) UG_THROW("LU::init: Passed operator is not a matrix operator.");
		return init_lu(m_spOperator->get_matrix());
	}
	virtual bool init(SmartPtr<ILinearOperator<vector_type> > J, const vector_type&) { return init(J); }
	bool init_lu(matrix_type& A)
	{
		const size_t nb = A.num_rows();
		m_size = nb * B;
		if (m_size == 0) return true;
		UG_COND_THROW(m_size > 4096, "LU: dense device LU limited to 4096 unknowns, use a CG base solver");
		const std::vector<int64_t>& rp = A.crs_rowptr(); const std::vector<int>& ci = A.crs_cols(); const std::vector<double>& va = A.crs_vals();
		const size_t n = m_size; const int BB = B * B;
		std::vector<double> a(n * n, 0.0); std::vector<int> piv(n, 0);
#define AA(i, j) a[(size_t)(i) * n + (j)]
		for (size_t r = 0; r < nb; ++r)
			for (int64_t p = rp[r]; p < rp[r + 1]; ++p)
				for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j) AA(r * B + i, (size_t)ci[p] * B + j) = va[p * BB + i + B * j];
		// LUDecomp with row interchange (no_lapack/lu_decomp.h:45-75)
		for (size_t k = 0; k < n; ++k) {
			size_t biggest = k;
			for (size_t j = k + 1; j < n; ++j) if (std::fabs(AA(biggest, k)) < std::fabs(AA(j, k))) biggest = j;
			if (biggest != k) for (size_t j = 0; j < n; ++j) std::swap(AA(k, j), AA(biggest, j));
			piv[k] = (int)biggest;
			if (std::fabs(AA(k, k)) < 1e-10) UG_THROW("ERROR in Matrix is singular");
			for (size_t i = k + 1; i < n; ++i) {
				AA(i, k) = AA(i, k) / AA(k, k);
				for (size_t j = k + 1; j < n; ++j) AA(i, j) = AA(i, j) - AA(i, k) * AA(k, j);
			}
		}
#undef AA
		free_dev();
		GPUManager::bump_generation();
		ug4b200_ctx* c = GPUManager::ctx();
		m_lu = (double*)GPUManager::alloc_bytes(sizeof(double) * n * n);
		m_piv = (int*)GPUManager::alloc_bytes(sizeof(int) * n);
		UG_GPU_CHECK(ug4b200_h2d(c, m_lu, a.data(), sizeof(double) * n * n));
		UG_GPU_CHECK(ug4b200_h2d(c, m_piv, piv.data(), sizeof(int) * n));
		UG_GPU_CHECK(ug4b200_sync(c));
		return true;
	}
	virtual bool apply(vector_type& u, const vector_type& f)
	{
		if (m_size == 0) return true;
		THROW_IF_NOT_EQUAL(f.len(), m_size);
		UG_GPU_CHECK(ug4b200_lu_apply(GPUManager::ctx(), (int)m_size, m_lu, m_piv, u.dev(), f.dev()));
		u.set_storage_type(PST_CONSISTENT);
		return true;
	}
	virtual bool apply_return_defect(vector_type& u, vector_type& f)
	{
		if (!apply(u, f)) return false;
		m_spOperator->apply_sub(f, u);
		return true;
	}
  protected:
	void free_dev() { GPUManager::free_bytes(m_lu); GPUManager::free_bytes(m_piv); m_lu = nullptr; m_piv = nullptr; }
	SmartPtr<matrix_operator_type> m_spOperator;
	size_t m_size = 0;
	double* m_lu = nullptr;
	int* m_piv = nullptr;
};

/// Small on-device CG as base solver (north_star: "coarse-grid solve done as a small
/// on-device CG"): one CTA runs the whole unpreconditioned CG (cg.h:103-242 with z = r)
template <typename TAlgebra>
class CoarseCG : public ILinearOperatorInverse<typename TAlgebra::vector_type> {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	typedef typename TAlgebra::matrix_type matrix_type;
	typedef MatrixOperator<matrix_type, vector_type> matrix_operator_type;
	enum { B = TAlgebra::blockSize };
	~CoarseCG() { if (m_work) GPUManager::release(m_work, m_n * 4); }
	virtual const char* name() const { return "CoarseCG"; }
	virtual bool supports_parallel() const { return false; }
	virtual bool init(SmartPtr<ILinearOperator<vector_type> > L)
	{
		m_spOperator = sp_cast_dynamic<matrix_operator_type>(L);
		if (!m_spOperator) UG_THROW("CoarseCG::init: Passed operator is not a matrix operator.");
		if (m_work) GPUManager::release(m_work, m_n * 4);
		GPUManager::bump_generation();
		m_n = m_spOperator->num_rows() * B;
		m_work = GPUManager::alloc(m_n * 4);
		m_spOperator->device();
		return true;
	}
	virtual bool init(SmartPtr<ILinearOperator<vector_type> > J, const vector_type&) { return init(J); }
	virtual bool apply(vector_type& u, const vector_type& f)
	{
		StdConvCheck<vector_type>* cc = dynamic_cast<StdConvCheck<vector_type>*>(this->convergence_check().get());
		const int maxSteps = cc ? cc->maximum_steps() : 1000;
		const number minDef = cc ? cc->minimum_defect() : 1e-30, red = cc ? cc->relative_reduction() : 1e-14;
		UG_GPU_CHECK(ug4b200_coarse_cg(GPUManager::ctx(), m_spOperator->device(), u.dev(), f.dev(), m_work, maxSteps, minDef, red));
		u.set_storage_type(PST_CONSISTENT);
		return true;
	}
	virtual bool apply_return_defect(vector_type& u, vector_type& f)
	{
		if (!apply(u, f)) return false;
		m_spOperator->apply_sub(f, u);
		return true;
	}
  protected:
	SmartPtr<matrix_operator_type> m_spOperator;
	size_t m_n = 0;
	double* m_work = nullptr;
};

} // namespace ug
