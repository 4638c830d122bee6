/*
 * oracle/solvers.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restatement of the solver-level control flow of ugcore's assembled-matrix
 * linear-solve path.  The reference classes themselves cannot be compiled here
 * (operator/convergence_check.h pulls lib_disc -> boost::mpl, absent), so the
 * loops are restated statement by statement; all arithmetic goes through a
 * Backend (either the port or the real compiled reference templates).
 * Paths below are relative to /root/reference/ugbase.
 *
 *   StdConvCheck     lib_algebra/operator/convergence_check_impl.h:85-169
 *   IPreconditioner  lib_algebra/operator/interface/preconditioner.h:192-348
 *   Jacobi           lib_algebra/operator/preconditioner/jacobi.h:155-300
 *   GaussSeidel*     lib_algebra/operator/preconditioner/gauss_seidel.h:114-380
 *   LU               lib_algebra/operator/linear_solver/lu.h:122-380
 *   CG               lib_algebra/operator/linear_solver/cg.h:103-242
 *   BiCGStab         lib_algebra/operator/linear_solver/bicgstab.h:112-383
 *   LinearSolver     lib_algebra/operator/linear_solver/linear_solver.h:114-196
 *   GMRES            lib_algebra/operator/linear_solver/gmres.h:104-278
 *   ILU              lib_algebra/operator/preconditioner/ilu.h:516-660 (kernels through the Backend:
 *                    the compiled reference's FactorizeILUSorted/Beta, invert_L, invert_U, or the port)
 *   GMG              lib_disc/operator/linear_operator/multi_grid_solver/
 *                    mg_solver_impl.hpp:174-275, 1685-1964, 1967-2136
 *   StdTransfer      lib_disc/operator/linear_operator/std_transfer_impl.h:738-740, 791-792
 */
#ifndef ORACLE_SOLVERS_H
#define ORACLE_SOLVERS_H
#include "backend.h"
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace oracle {

/// 0: the reference's sequential reductions (default); 1: pairwise tree (measurement aid, see solvers.cpp)
void set_reduction_mode(int mode);
int reduction_mode();


typedef std::unique_ptr<Vec> VecP;

struct StdConvCheck {
	int maxSteps = 100;
	double minDefect = 1e-12, relReduction = 1e-6;
	double initialDefect = 0, currentDefect = 0, lastDefect = 0, ratesProduct = 1;
	int currentStep = 0;
	std::vector<double> history; // defect after start and after every update

	void start_defect(double d);
	void update_defect(double d);
	bool iteration_ended() const;
	bool post() const;
	double defect() const { return currentDefect; }
	double reduction() const { return currentDefect / initialDefect; }
	int step() const { return currentStep; }
	static bool is_valid_number(double v);
};

struct LinearIterator {
	Backend& bk;
	explicit LinearIterator(Backend& b) : bk(b) {}
	virtual ~LinearIterator() {}
	virtual const char* name() const = 0;
	virtual bool init(const Mat& A) = 0;
	virtual bool apply(Vec& c, const Vec& d) = 0;
	virtual bool apply_update_defect(Vec& c, Vec& d) = 0;
	virtual LinearIterator* clone() const = 0;
};

struct Preconditioner : LinearIterator {
	const Mat* A = nullptr;
	double damping = 1.0; // ConstantDamping, damping.h:100-127
	bool inited = false;
	explicit Preconditioner(Backend& b) : LinearIterator(b) {}
	bool init(const Mat& A_) override;
	bool apply(Vec& c, const Vec& d) override;
	bool apply_update_defect(Vec& c, Vec& d) override;
	virtual bool preprocess() = 0;
	virtual bool step(Vec& c, const Vec& d) = 0;
};

struct Jacobi : Preconditioner {
	bool block = true; // set_block, jacobi.h:210-213 (default m_bBlock = true)
	std::unique_ptr<DiagInv> diagInv;
	explicit Jacobi(Backend& b, double damp = 1.0) : Preconditioner(b) { damping = damp; }
	const char* name() const override { return "Jacobi"; }
	bool preprocess() override;
	bool step(Vec& c, const Vec& d) override;
	bool apply(Vec& c, const Vec& d) override; // jacobi.h:256-300: damping folded into diagInv
	LinearIterator* clone() const override { Jacobi* j = new Jacobi(bk, damping); j->block = block; return j; }
};

struct GaussSeidel : Preconditioner {
	enum Kind { FORWARD, BACKWARD, SYMMETRIC } kind;
	double relax = 1.0;
	GaussSeidel(Backend& b, Kind k) : Preconditioner(b), kind(k) {}
	const char* name() const override
	{ return kind == FORWARD ? "Gauss-Seidel" : kind == BACKWARD ? "Backward Gauss-Seidel" : "Symmetric Gauss-Seidel"; }
	bool preprocess() override { return true; }
	bool step(Vec& c, const Vec& d) override;
	LinearIterator* clone() const override
	{ GaussSeidel* g = new GaussSeidel(bk, kind); g->relax = relax; g->damping = damping; return g; }
};

struct ILU : Preconditioner {
	double beta = 0.0, sortEps = 1e-50, invEps = 1e-8; // ilu.h:352-355
	std::unique_ptr<Mat> factors;
	VecP h;
	explicit ILU(Backend& b, double beta_ = 0.0) : Preconditioner(b), beta(beta_) {}
	const char* name() const override { return "ILU"; }
	bool preprocess() override;                        // ilu.h:516-588 (serial, no ordering algorithm)
	bool step(Vec& c, const Vec& d) override;          // ilu.h:591-599, 648-652
	LinearIterator* clone() const override
	{ ILU* g = new ILU(bk, beta); g->sortEps = sortEps; g->invEps = invEps; g->damping = damping; return g; }
};

struct InverseOperator {
	Backend& bk;
	const Mat* A = nullptr;
	StdConvCheck conv;
	explicit InverseOperator(Backend& b) : bk(b) {}
	virtual ~InverseOperator() {}
	virtual const char* name() const = 0;
	virtual bool init(const Mat& A_) { A = &A_; return true; }
	virtual bool apply_return_defect(Vec& x, Vec& b) = 0;
	// preconditioned_linear_operator_inverse.h:152-160
	virtual bool apply(Vec& x, const Vec& b);
};

struct LU : InverseOperator {
	std::unique_ptr<DenseLU> lu;
	explicit LU(Backend& b) : InverseOperator(b) {}
	const char* name() const override { return "LU"; }
	bool init(const Mat& A_) override;
	bool apply(Vec& x, const Vec& b) override;
	bool apply_return_defect(Vec& x, Vec& b) override;
};

struct PrecondInverse : InverseOperator {
	std::unique_ptr<LinearIterator> precond;
	explicit PrecondInverse(Backend& b) : InverseOperator(b) {}
	bool init(const Mat& A_) override
	{
		A = &A_;
		if (precond && !precond->init(A_)) return false;
		return true;
	}
};

struct CG : PrecondInverse {
	explicit CG(Backend& b) : PrecondInverse(b) {}
	const char* name() const override { return "CG"; }
	bool apply_return_defect(Vec& x, Vec& b) override;
};

struct BiCGStab : PrecondInverse {
	int numRestarts = 0;
	double minOrtho = 0.0;
	explicit BiCGStab(Backend& b) : PrecondInverse(b) {}
	const char* name() const override { return "BiCGStab"; }
	bool apply_return_defect(Vec& x, Vec& b) override;
};

struct GMRES : PrecondInverse {
	size_t restart = 30;
	explicit GMRES(Backend& b) : PrecondInverse(b) {}
	const char* name() const override { return "GMRES"; }
	bool apply_return_defect(Vec& x, Vec& b) override;
};

struct LinearSolver : PrecondInverse {
	explicit LinearSolver(Backend& b) : PrecondInverse(b) {}
	const char* name() const override { return "Iterative Linear Solver"; }
	bool apply_return_defect(Vec& x, Vec& b) override;
};

// AssembledMultiGridCycle on a fully refined hierarchy (surface == top level, so
// the surface<->level index maps are identities and m_LocalFullRefLevel == topLev).
struct GMG : LinearIterator {
	enum { V_CYCLE = 1, W_CYCLE = 2, F_CYCLE = -1 };
	struct LevData {
		const Mat* A = nullptr;
		const Mat* P = nullptr; // level-1 -> level
		const Mat* R = nullptr; // level -> level-1
		const Mat* S = nullptr; // matrix the smoothers are initialised with (nullptr: A).  ugcore's parallel
		                        // Gauss-Seidel smooths with its own matrix m_A (gauss_seidel.h:134-142, 225-231):
		                        // in global terms the level matrix without the couplings between DoFs of
		                        // different h-masters — tests hand that matrix in here
		VecP sc, sd, st;
		std::unique_ptr<LinearIterator> pre, post;
	};
	int baseLev = 0, topLev = 0;
	int cycleType = V_CYCLE;
	int numPreSmooth = 2, numPostSmooth = 2;
	double dampProl = 1.0, dampRes = 1.0; // std_transfer.h m_dampProl / m_dampRes
	std::unique_ptr<LinearIterator> smootherProto;
	std::unique_ptr<InverseOperator> baseSolver;
	std::vector<LevData> lev; // index = level - baseLev
	double damping = 1.0;
	// surface index of every top-level index (vSurfLevelMap of a fully refined grid whose surface and level DoF
	// distributions number the DoFs differently, mg_solver_impl.hpp:1344-1369); empty = identity
	std::vector<int> surfMap;

	explicit GMG(Backend& b) : LinearIterator(b) {}
	const char* name() const override { return "Geometric MultiGrid"; }
	void set_level(int level, const Mat* A, const Mat* P, const Mat* R);
	void set_level_smoother_matrix(int level, const Mat* S);
	LevData& L(int l) { return lev[l - baseLev]; }
	bool init(const Mat& A_) override;
	bool apply(Vec& c, const Vec& d) override;
	bool apply_update_defect(Vec& c, Vec& d) override;
	LinearIterator* clone() const override { return nullptr; } // not needed by the oracle
  private:
	const Mat* surfaceMat = nullptr;
	void lmgc(int l, int cycle);
	void presmooth_and_restriction(int l);
	void prolongation_and_postsmooth(int l);
	void base_solve(int l);
};

} // namespace oracle
#endif
