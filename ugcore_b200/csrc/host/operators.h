// operators.h — ugcore's operator interfaces for the GPU algebra, signatures kept verbatim.
//
//   ILinearOperator            ugbase/lib_algebra/operator/interface/linear_operator.h:78-143
//   MatrixOperator             ugbase/lib_algebra/operator/interface/matrix_operator.h:46-75
//   ILinearIterator            ugbase/lib_algebra/operator/interface/linear_iterator.h:79-198
//   IPreconditioner            ugbase/lib_algebra/operator/interface/preconditioner.h:99-388
//   ILinearOperatorInverse     ugbase/lib_algebra/operator/interface/linear_operator_inverse.h:78-242
//   IPreconditionedLinearOperatorInverse
//                              ugbase/lib_algebra/operator/interface/preconditioned_linear_operator_inverse.h:59-237
//   IConvergenceCheck/StdConvCheck  ugbase/lib_algebra/operator/convergence_check.h, _impl.h:85-169
//   IDamping/ConstantDamping   ugbase/lib_algebra/operator/damping.h:100-127
//
// Inside a ugcore build (UG4B200_WITH_UGCORE) the four interface headers that compile without boost are the REAL
// ones — ILinearOperator, MatrixOperator, IDamping / ConstantDamping, ILinearIterator, IPreconditioner,
// IVectorDebugWriter — and only the three classes whose ugcore headers pull boost::mpl through
// lib_disc/domain_traits.h (convergence_check.h, linear_operator_inverse.h,
// preconditioned_linear_operator_inverse.h) keep their restatement below.  tests/boundary/ builds that
// configuration against /root/reference.
#pragma once
#include "gpu_sparsematrix.h"
#include <limits>

#ifdef UG4B200_WITH_UGCORE
#include "lib_algebra/operator/interface/linear_operator.h"
#include "lib_algebra/operator/interface/matrix_operator.h"
#include "lib_algebra/operator/interface/linear_iterator.h"
#include "lib_algebra/operator/interface/preconditioner.h"
#include "lib_algebra/operator/debug_writer.h"
#endif

namespace ug {

#ifndef UG4B200_WITH_UGCORE
template <typename X, typename Y = X>
class ILinearOperator {
  public:
	typedef X domain_function_type;
	typedef Y codomain_function_type;
	virtual void init(const X& u) = 0;
	virtual void init() = 0;
	virtual void apply(Y& f, const X& u) = 0;
	virtual void apply_sub(Y& f, const X& u) = 0;
	virtual ~ILinearOperator() {}
};

template <typename M, typename X, typename Y = X>
class MatrixOperator : public virtual ILinearOperator<X, Y>, public M {
  public:
	typedef M matrix_type;
	virtual void init(const X&) {}
	virtual void init() {}
	virtual void apply(Y& f, const X& u) { matrix_type::apply(f, u); }
	virtual void apply_sub(Y& f, const X& u) { matrix_type::matmul_minus(f, u); }
	virtual M& get_matrix() { return *this; }
};

// ---- damping ----
template <typename X, typename Y = X>
class IDamping {
  public:
	virtual number damping(const Y& c, const X& d, ConstSmartPtr<ILinearOperator<Y, X> > spLinOp) const = 0;
	virtual number damping() const = 0;
	virtual bool constant_damping() const = 0;
	virtual ~IDamping() {}
};
template <typename X, typename Y = X>
class ConstantDamping : public IDamping<X, Y> {
  public:
	explicit ConstantDamping(number factor) : m_factor(factor) {}
	virtual number damping(const Y&, const X&, ConstSmartPtr<ILinearOperator<Y, X> >) const { return m_factor; }
	virtual number damping() const { return m_factor; }
	virtual bool constant_damping() const { return true; }
  protected:
	number m_factor;
};

template <typename X, typename Y = X>
class ILinearIterator {
  public:
	ILinearIterator() { set_damp(1.0); }
	virtual ~ILinearIterator() {}
	virtual const char* name() const = 0;
	virtual bool supports_parallel() const = 0;
	virtual bool init(SmartPtr<ILinearOperator<Y, X> > J, const Y& u) = 0;
	virtual bool init(SmartPtr<ILinearOperator<Y, X> > L) = 0;
	virtual bool apply(Y& c, const X& d) = 0;
	virtual bool apply_update_defect(Y& c, X& d) = 0;
	virtual SmartPtr<ILinearIterator<X, Y> > clone() = 0;
	void set_damp(SmartPtr<IDamping<X, Y> > spScaling) { m_spDamping = spScaling; }
	void set_damp(number factor) { m_spDamping = SmartPtr<IDamping<X, Y> >(new ConstantDamping<X, Y>(factor)); }
	SmartPtr<IDamping<X, Y> > damping() { return m_spDamping; }
  protected:
	SmartPtr<IDamping<X, Y> > m_spDamping;
};

template <typename TAlgebra>
class IPreconditioner : public ILinearIterator<typename TAlgebra::vector_type> {
  public:
	typedef TAlgebra algebra_type;
	typedef typename TAlgebra::vector_type vector_type;
	typedef typename TAlgebra::matrix_type matrix_type;
	typedef MatrixOperator<matrix_type, vector_type> matrix_operator_type;
	using ILinearIterator<vector_type>::damping;

	IPreconditioner() : m_bInit(false) {}
	virtual const char* name() const = 0;

	virtual bool init(SmartPtr<ILinearOperator<vector_type> > J, const vector_type&)
	{
		SmartPtr<matrix_operator_type> pOp = sp_cast_dynamic<matrix_operator_type>(J);
		if (!pOp) UG_THROW(name() << "::init': Passed Operator is not based on matrix. This Preconditioner can only handle matrix-based operators.");
		return init(pOp);
	}
	virtual bool init(SmartPtr<ILinearOperator<vector_type> > L)
	{
		SmartPtr<matrix_operator_type> pOp = sp_cast_dynamic<matrix_operator_type>(L);
		if (!pOp) UG_THROW(name() << "::init': Passed Operator is not based on matrix. This Preconditioner can only handle matrix-based operators.");
		return init(pOp);
	}
	bool init(SmartPtr<matrix_operator_type> Op)
	{
		m_spApproxOperator = Op; m_spDefectOperator = Op;
		if (!m_spApproxOperator) UG_THROW(name() << "::init': Passed Operator is invalid.");
		if (!preprocess(m_spApproxOperator)) return false;
		m_bInit = true;
		return true;
	}
	virtual bool apply(vector_type& c, const vector_type& d)
	{
		if (!m_bInit) return false;
		if (d.layouts() && !d.has_storage_type(PST_ADDITIVE))
			UG_THROW(name() << "::apply: Wrong parallel storage format. Defect must be additive.");
		THROW_IF_NOT_EQUAL(c.size(), d.size());
		THROW_IF_NOT_EQUAL(c.size(), m_spApproxOperator->num_rows());
		if (!step(m_spApproxOperator, c, d)) return false;
		const number kappa = damping()->damping(c, d, m_spApproxOperator);
		if (kappa != 1.0) c *= kappa;
		if (c.layouts() && !c.change_storage_type(PST_CONSISTENT))
			UG_THROW(name() << "::apply': Cannot change parallel storage type of correction to consistent.");
		return true;
	}
	virtual bool apply_update_defect(vector_type& c, vector_type& d)
	{
		if (!apply(c, d)) return false;
		m_spDefectOperator->apply_sub(d, c);
		return true;
	}
	SmartPtr<matrix_operator_type> approx_operator() { return m_spApproxOperator; }
	SmartPtr<ILinearOperator<vector_type> > defect_operator() { return m_spDefectOperator; }

  protected:
	virtual bool preprocess(SmartPtr<matrix_operator_type> pOp) = 0;
	virtual bool step(SmartPtr<matrix_operator_type> pOp, vector_type& c, const vector_type& d) = 0;
	virtual bool postprocess() = 0;

	SmartPtr<ILinearOperator<vector_type> > m_spDefectOperator;
	SmartPtr<matrix_operator_type> m_spApproxOperator;
	bool m_bInit;
};

#endif // !UG4B200_WITH_UGCORE

// ---- convergence check ----
template <typename TVector>
class IConvergenceCheck {
  public:
	virtual void start_defect(number defect) = 0;
	virtual void start(const TVector& d) = 0;
	virtual void update_defect(number defect) = 0;
	virtual void update(const TVector& d) = 0;
	virtual bool iteration_ended() = 0;
	virtual bool post() = 0;
	virtual number defect() const = 0;
	virtual int step() const = 0;
	virtual number reduction() const = 0;
	virtual number rate() const = 0;
	virtual number avg_rate() const = 0;
	virtual ~IConvergenceCheck() {}
};

template <typename TVector>
class StdConvCheck : public IConvergenceCheck<TVector> {
  public:
	StdConvCheck() : StdConvCheck(100, 1e-12, 1e-12, true) {}
	StdConvCheck(int maxSteps, number minDefect, number relReduction, bool verbose = false)
	    : m_initialDefect(0.0), m_currentDefect(0.0), m_lastDefect(0.0), m_currentStep(0), m_ratesProduct(1),
	      m_maxSteps(maxSteps), m_minDefect(minDefect), m_relReduction(relReduction), m_verbose(verbose) {}

	void set_maximum_steps(int maxSteps) { m_maxSteps = maxSteps; }
	void set_minimum_defect(number minDefect) { m_minDefect = minDefect; }
	void set_reduction(number relReduction) { m_relReduction = relReduction; }
	void set_verbose(bool level) { m_verbose = level; }
	int maximum_steps() const { return m_maxSteps; }
	number minimum_defect() const { return m_minDefect; }
	number relative_reduction() const { return m_relReduction; }

	void start_defect(number initialDefect)
	{
		_defects.clear();
		m_initialDefect = initialDefect; m_currentDefect = m_initialDefect;
		m_currentStep = 0; m_ratesProduct = 1;
		_defects.push_back(initialDefect);
	}
	void start(const TVector& d) { start_defect(d.norm()); }
	void update_defect(number newDefect)
	{
		m_lastDefect = m_currentDefect; m_currentDefect = newDefect; m_currentStep++;
		m_ratesProduct *= newDefect / m_lastDefect;
		_defects.push_back(newDefect);
	}
	void update(const TVector& d) { update_defect(d.norm()); }
	bool iteration_ended()
	{
		if (!is_valid_number(m_currentDefect)) return true;
		if (step() >= m_maxSteps) return true;
		if (defect() < m_minDefect) return true;
		if (reduction() < m_relReduction) return true;
		return false;
	}
	bool post()
	{
		bool success = false;
		if (defect() < m_minDefect) success = true;
		if (reduction() < m_relReduction) success = true;
		return success;
	}
	number reduction() const { return m_currentDefect / m_initialDefect; }
	number rate() const { return m_currentDefect / m_lastDefect; }
	number avg_rate() const { return std::pow((number)m_ratesProduct, (number)1.0 / (number)m_currentStep); }
	number defect() const { return m_currentDefect; }
	number previous_defect() const { return m_lastDefect; }
	int step() const { return m_currentStep; }
	/// defect history; entry 0 is the start defect (the reference records updates only when verbose)
	const std::vector<number>& get_defects() const { return _defects; }

	/// adopt the outcome of a device-side iteration (ug4b200_conv_state + history)
	void adopt_device_state(const ug4b200_conv_state& s, const std::vector<number>& history)
	{
		m_initialDefect = s.initial_defect; m_currentDefect = s.current_defect; m_lastDefect = s.last_defect;
		m_currentStep = s.step; _defects = history;
		m_ratesProduct = (m_initialDefect != 0.0) ? m_currentDefect / m_initialDefect : 1.0;
	}
	static bool is_valid_number(number value)
	{
		if (value == 0.0) return true;
		return value >= std::numeric_limits<number>::min() && value <= std::numeric_limits<number>::max() &&
		       value == value && value >= 0.0;
	}

  protected:
	number m_initialDefect, m_currentDefect, m_lastDefect;
	int m_currentStep;
	number m_ratesProduct;
	int m_maxSteps;
	number m_minDefect, m_relReduction;
	bool m_verbose;
	std::vector<number> _defects;
};

/// ugbase/lib_algebra/operator/debug_writer.h: IVectorDebugWriter — receives vectors by name while a solver runs
/// (CG_Residual_iterNNN.vec …); concrete writer: ConnectionViewerVectorWriter (matrix_io.h)
#ifndef UG4B200_WITH_UGCORE
template <typename TVector>
class IVectorDebugWriter {
  public:
	virtual ~IVectorDebugWriter() {}
	virtual void write_vector(const TVector& vec, const char* name) = 0;
};
#endif

template <typename X, typename Y = X>
class ILinearOperatorInverse {
  public:
	ILinearOperatorInverse() : m_spConvCheck(new StdConvCheck<X>(100, 1e-12, 1e-12, true)) {}
	virtual ~ILinearOperatorInverse() {}
	/// VectorDebugWritingObject (debug_writer.h:262-330): set_debug / vector_debug_writer_valid / write_debug
	void set_debug(SmartPtr<IVectorDebugWriter<X> > spDebugWriter) { m_spVectorDebugWriter = spDebugWriter; }
	bool vector_debug_writer_valid() const { return (bool)m_spVectorDebugWriter; }
	void write_debug(const X& vec, const std::string& name) { if (m_spVectorDebugWriter) m_spVectorDebugWriter->write_vector(vec, name.c_str()); }
	virtual const char* name() const = 0;
	virtual bool supports_parallel() const = 0;
	virtual bool init(SmartPtr<ILinearOperator<Y, X> > L) { GPUManager::bump_generation(); m_spLinearOperator = L; return true; }
	virtual bool init(SmartPtr<ILinearOperator<Y, X> > J, const Y&) { GPUManager::bump_generation(); m_spLinearOperator = J; return true; }
	virtual bool apply(Y& u, const X& f) = 0;
	virtual bool apply_return_defect(Y& u, X& f) = 0;
	number defect() const { return convergence_check()->defect(); }
	int step() const { return convergence_check()->step(); }
	number reduction() const { return convergence_check()->reduction(); }
	void set_convergence_check(SmartPtr<IConvergenceCheck<X> > spConvCheck) { m_spConvCheck = spConvCheck; }
	SmartPtr<IConvergenceCheck<X> > convergence_check() { return m_spConvCheck; }
	ConstSmartPtr<IConvergenceCheck<X> > convergence_check() const { return m_spConvCheck; }
	SmartPtr<ILinearOperator<Y, X> > linear_operator() { return m_spLinearOperator; }
  protected:
	SmartPtr<ILinearOperator<Y, X> > m_spLinearOperator;
	SmartPtr<IConvergenceCheck<X> > m_spConvCheck;
	SmartPtr<IVectorDebugWriter<X> > m_spVectorDebugWriter;
};

template <typename X>
class IPreconditionedLinearOperatorInverse : public ILinearOperatorInverse<X> {
  public:
	typedef ILinearOperatorInverse<X> base_type;
	using base_type::linear_operator;
	using base_type::name;
	IPreconditionedLinearOperatorInverse() {}
	explicit IPreconditionedLinearOperatorInverse(SmartPtr<ILinearIterator<X, X> > spPrecond) : m_spPrecond(spPrecond) {}
	void set_preconditioner(SmartPtr<ILinearIterator<X, X> > spPrecond) { GPUManager::bump_generation(); m_spPrecond = spPrecond; }
	SmartPtr<ILinearIterator<X, X> > preconditioner() { return m_spPrecond; }
	virtual bool supports_parallel() const { return m_spPrecond ? m_spPrecond->supports_parallel() : true; }
	virtual bool init(SmartPtr<ILinearOperator<X, X> > J, const X& u)
	{
		if (!base_type::init(J, u)) return false;
		if (m_spPrecond && !m_spPrecond->init(J, u)) UG_THROW(name() << "::init: Cannot init Preconditioner Operator for Operator J.");
		return true;
	}
	virtual bool init(SmartPtr<ILinearOperator<X, X> > L)
	{
		if (!base_type::init(L)) return false;
		if (m_spPrecond && !m_spPrecond->init(L)) UG_THROW(name() << "::prepare: Cannot init Preconditioner Operator for Operator L.");
		return true;
	}
	/// preconditioned_linear_operator_inverse.h:152-160: b is cloned, the defect is dropped
	virtual bool apply(X& x, const X& b)
	{
		SmartPtr<X> spB = b.clone();
		return this->apply_return_defect(x, *spB);
	}
  protected:
	SmartPtr<ILinearIterator<X, X> > m_spPrecond;
};

} // namespace ug
