// gpu_vector.h — GPUVector<T>: device-resident Vector<T> with a lazily mirrored host copy.
//
// Mirrors ugbase/lib_algebra/cpu_algebra/vector.h:53-230 (API used by GridFunction and the
// solvers) and, for partitioned runs, the ParallelVector wrapper
// (ugbase/lib_algebra/parallelization/parallel_vector.h:58-198, parallel_vector_impl.h
// :115-218 storage-type changes, :269-294 norm, :323-379 dotprod).  Shape of the drop-in:
// the legacy GPUVector (ugbase/lib_algebra/gpu_algebra/gpuvector.h:50-219) with its
// ON_CPU/ON_GPU dirty bits; arithmetic never touches the host.
#pragma once
#include "ug_base.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace ug {

/// Horizontal layouts + communicator of one level (AlgebraLayouts, algebra_layouts.h:47-150)
class GPUAlgebraLayouts {
  public:
	GPUAlgebraLayouts(int nneigh, const int* neighRank, const int64_t* neighPtr, const int* indices, int64_t nlocal, int myRank = -1)
	    : m_rank(myRank), m_nlocal(nlocal), m_neighRank(neighRank, neighRank + nneigh), m_neighPtr(neighPtr, neighPtr + nneigh + 1),
	      m_indices(indices, indices + (nneigh ? neighPtr[nneigh] : 0))
	{
		UG_GPU_CHECK(ug4b200_interface_create(GPUManager::ctx(), nneigh, neighRank, neighPtr, indices, nlocal, &m_iface));
		// peer-window transport: look up where the neighbours receive (they publish it when they
		// create the same interface; every rank builds its layouts in the same order)
		UG_GPU_CHECK(ug4b200_interface_commit(GPUManager::ctx(), m_iface));
	}
	~GPUAlgebraLayouts() { if (m_iface && GPUManager::ctx_or_null()) ug4b200_interface_destroy(GPUManager::ctx_or_null(), m_iface); }
	ug4b200_interface* iface() const { return m_iface; }
	int64_t num_local() const { return m_nlocal; }
	/// h-slave DoFs of this rank: shared with a rank of lower number (the h-master of a DoF is the lowest
	/// rank that holds a copy; IndexLayout slave(), algebra_layouts.h:47-150), ascending, unique.
	/// Needs the rank the layouts were created for.
	std::vector<int> slave_indices() const
	{
		if (m_rank < 0) UG_THROW("GPUAlgebraLayouts::slave_indices: rank of the layouts unknown");
		std::vector<char> slave((size_t)m_nlocal, 0);
		for (size_t p = 0; p < m_neighRank.size(); ++p)
			if (m_neighRank[p] < m_rank)
				for (int64_t k = m_neighPtr[p]; k < m_neighPtr[p + 1]; ++k) slave[(size_t)m_indices[(size_t)k]] = 1;
		std::vector<int> out;
		for (int64_t i = 0; i < m_nlocal; ++i) if (slave[(size_t)i]) out.push_back((int)i);
		return out;
	}
  private:
	ug4b200_interface* m_iface = nullptr;
	int m_rank = -1;
	int64_t m_nlocal = 0;
	std::vector<int> m_neighRank;
	std::vector<int64_t> m_neighPtr;
	std::vector<int> m_indices;
};

template <typename TValueType>
class GPUVector {
  public:
	typedef TValueType value_type;
	typedef GPUVector<TValueType> vector_type, this_type;
	enum { blockSize = block_traits<TValueType>::static_size };

	GPUVector() {}
	explicit GPUVector(size_t n) { create(n); }
	GPUVector(const GPUVector& v) { *this = v; }
	virtual ~GPUVector() { destroy(); }

	// ---- size / creation (vector.h:75-118) ----
	size_t size() const { return m_size; }
	size_t len() const { return m_size * blockSize; }
	void create(size_t n) { destroy(); m_size = n; m_dev = GPUManager::alloc(len()); m_hostValid = false; m_devValid = true; set(0.0); }
	void resize(size_t n, bool bCopyValues = true)
	{
		if (n == m_size) return;
		GPUVector tmp(n);
		if (bCopyValues && m_size) {
			const size_t c = (n < m_size ? n : m_size) * blockSize;
			UG_GPU_CHECK(ug4b200_vec_copy(GPUManager::ctx(), c, tmp.dev(), dev()));
		}
		swap(tmp);
	}
	void resize_sloppy(size_t n, bool bCopyValues = true) { resize(n, bCopyValues); }
	void resize_exactly(size_t n, bool bCopyValues = true) { resize(n, bCopyValues); }

	virtual SmartPtr<this_type> clone() const { SmartPtr<this_type> v(new this_type(*this)); return v; }
	virtual SmartPtr<this_type> clone_without_values() const
	{
		SmartPtr<this_type> v(new this_type(m_size));
		v->m_layouts = m_layouts; v->m_type = PST_UNDEFINED;
		return v;
	}

	void create(const this_type& v) { create(v.size()); }                              // vector.h:84
	void reserve_sloppy(size_t, bool = true) {}                                        // capacity == size on the device
	void reserve_exactly(size_t, bool) {}
	void reserve(size_t, bool = true) {}
	size_t capacity() const { return m_size; }
	/// values[i] = urand(from, to) in index order with the C library's rand(), like the reference
	/// (vector_impl.h:91-96, common/math/misc/math_util_impl.hpp:64-74): same seed, same vector
	void set_random(double from, double to)
	{
		std::vector<double> h(len());
		for (size_t i = 0; i < h.size(); ++i) {
			long t = std::rand();
			if (t == RAND_MAX) t -= 1;
			h[i] = from + (double)((to - from) * ((double)t / (double)RAND_MAX));
		}
		if (!h.empty()) assign_from_host(h.data());
	}
	/// max_i BlockMaxNorm(values[i]) (vector_impl.h:332-338); not on the solve path: evaluated on the host mirror
	double maxnorm() const
	{
		const_cast<this_type*>(this)->to_host();
		const double* h = m_size ? reinterpret_cast<const double*>(&m_host[0]) : nullptr;
		double d = 0;
		for (size_t i = 0; i < len(); ++i) d = std::max(d, std::fabs(h[i]));
		return d;
	}
	/// element-wise access by index lists (vector.h:158-160; assembly side: host mirror)
	void add(const value_type* u, const size_t* indices, size_t nr) { for (size_t i = 0; i < nr; ++i) add_block((*this)[indices[i]], u[i]); }
	void set(const value_type* u, const size_t* indices, size_t nr) { for (size_t i = 0; i < nr; ++i) (*this)[indices[i]] = u[i]; }
	void get(value_type* u, const size_t* indices, size_t nr) const { for (size_t i = 0; i < nr; ++i) u[i] = (*this)[indices[i]]; }
	/// local (element) vectors: V offers size(), index(i), operator[](i) (vector.h:153-155)
	template <typename V> void add(const V& u) { for (size_t i = 0; i < u.size(); ++i) add_block((*this)[u.index(i)], u[i]); }
	template <typename V> void set(const V& u) { for (size_t i = 0; i < u.size(); ++i) (*this)[u.index(i)] = u[i]; }
	template <typename V> void get(V& u) const { for (size_t i = 0; i < u.size(); ++i) u[i] = (*this)[u.index(i)]; }

	// ---- host element access (assembly, Dirichlet adjust, output): forces a D2H mirror ----
	value_type& operator[](size_t i) { to_host(); m_devValid = false; return m_host[i]; }
	const value_type& operator[](size_t i) const { const_cast<this_type*>(this)->to_host(); return m_host[i]; }

	// ---- device access: the hot path only uses these ----
	double* dev() { to_device(); m_hostValid = false; return m_dev; }
	const double* dev() const { const_cast<this_type*>(this)->to_device(); return m_dev; }

	void assign_from_host(const double* h)
	{
		UG_GPU_CHECK(ug4b200_h2d(GPUManager::ctx(), m_dev, h, len() * sizeof(double)));
		UG_GPU_CHECK(ug4b200_sync(GPUManager::ctx()));
		m_devValid = true; m_hostValid = false;
	}
	void copy_to_host(double* h) const { UG_GPU_CHECK(ug4b200_d2h(GPUManager::ctx(), h, dev(), len() * sizeof(double))); }

	// ---- arithmetic (vector.h:124-176) ----
	void set(double d) { UG_GPU_CHECK(ug4b200_vec_set(GPUManager::ctx(), len(), dev_w(), d)); }
	double operator=(double d) { set(d); return d; }
	this_type& operator=(const this_type& v)
	{
		if (this == &v) return *this;
		if (m_size != v.m_size) { destroy(); m_size = v.m_size; m_dev = GPUManager::alloc(len()); }
		UG_GPU_CHECK(ug4b200_vec_copy(GPUManager::ctx(), len(), dev_w(), v.dev()));
		m_layouts = v.m_layouts; m_type = v.m_type;
		return *this;
	}
	this_type& operator+=(const this_type& v)
	{
		check_size(v);
		UG_GPU_CHECK(ug4b200_vec_add(GPUManager::ctx(), len(), dev(), v.dev()));
		m_type &= v.m_type; // parallel_vector_impl.h:79-90: only common storage types survive
		return *this;
	}
	this_type& operator-=(const this_type& v)
	{
		check_size(v);
		UG_GPU_CHECK(ug4b200_vec_sub(GPUManager::ctx(), len(), dev(), v.dev()));
		m_type &= v.m_type;
		return *this;
	}
	this_type& operator*=(const number& a) { UG_GPU_CHECK(ug4b200_vec_scale(GPUManager::ctx(), len(), dev(), a)); return *this; }

	/// Vector::dotprod (vector_impl.h:72-79) / ParallelVector::dotprod (parallel_vector_impl.h:323-379)
	double dotprod(const this_type& w)
	{
		check_size(w);
		if (m_layouts) {
			if (has_storage_type(PST_UNDEFINED) || w.has_storage_type(PST_UNDEFINED))
				UG_THROW("ERROR in ParallelVector::dotprod(): No parallel Storage type given.");
			bool check = (has_storage_type(PST_ADDITIVE) && w.has_storage_type(PST_CONSISTENT)) ||
			             (has_storage_type(PST_CONSISTENT) && w.has_storage_type(PST_ADDITIVE)) ||
			             (has_storage_type(PST_UNIQUE) && w.has_storage_type(PST_UNIQUE));
			if (!check) {
				if (has_storage_type(PST_UNIQUE) && w.has_storage_type(PST_ADDITIVE)) change_storage_type(PST_CONSISTENT);
				else change_storage_type(PST_UNIQUE);
			}
		}
		ug4b200_ctx* c = GPUManager::ctx();
		double r = 0.0;
		if (!m_layouts) { UG_GPU_CHECK(ug4b200_vec_dot(c, len(), dev(), w.dev(), &r)); return r; }
		double* slot = scalar_slot();
		ug4b200_fin fin{UG4B200_FIN_STORE, slot, nullptr, nullptr, nullptr};
		UG_GPU_CHECK(ug4b200_vec_dot_ds(c, len(), dev(), w.dev(), fin));
		UG_GPU_CHECK(ug4b200_allreduce_sum(c, slot, 1));
		UG_GPU_CHECK(ug4b200_d2h(c, &r, slot, sizeof(double)));
		return r;
	}
	/// Vector::norm (vector_impl.h:323-329) / ParallelVector::norm (parallel_vector_impl.h:269-294)
	double norm() const
	{
		ug4b200_ctx* c = GPUManager::ctx();
		double r = 0.0;
		if (!m_layouts) { UG_GPU_CHECK(ug4b200_vec_norm(c, len(), dev(), &r)); return r; }
		this_type* self = const_cast<this_type*>(this);
		if (!self->change_storage_type(PST_UNIQUE)) UG_THROW("ParallelVector::norm(): Cannot change ParallelStorageType to unique.");
		double* slot = scalar_slot();
		ug4b200_fin fin{UG4B200_FIN_STORE, slot, nullptr, nullptr, nullptr};
		UG_GPU_CHECK(ug4b200_vec_dot_ds(c, len(), dev(), dev(), fin));
		UG_GPU_CHECK(ug4b200_allreduce_sum(c, slot, 1));
		UG_GPU_CHECK(ug4b200_d2h(c, &r, slot, sizeof(double)));
		return std::sqrt(r);
	}

	// ---- parallel storage type (parallel_vector.h:115-160) ----
	void set_layouts(SmartPtr<GPUAlgebraLayouts> l) { m_layouts = l; }
	SmartPtr<GPUAlgebraLayouts> layouts() const { return m_layouts; }
	void set_storage_type(unsigned type) { m_type = type; }
	void add_storage_type(unsigned type) { m_type |= type; }
	bool has_storage_type(unsigned type) const { return type == PST_UNDEFINED ? m_type == PST_UNDEFINED : (m_type & type) == type; }
	unsigned get_storage_mask() const { return m_type; }
	bool change_storage_type(ParallelStorageType type)
	{
		if (!m_layouts) { m_type = type == PST_UNIQUE ? (PST_ADDITIVE | PST_UNIQUE) : type; return true; }
		if (has_storage_type(PST_UNDEFINED))
			UG_THROW("ParallelVector::change_storage_type: Trying to change storage type of a vector that has type PST_UNDEFINED.");
		if (has_storage_type(type)) return true;
		ug4b200_ctx* c = GPUManager::ctx();
		ug4b200_interface* I = m_layouts->iface();
		switch (type) {
			case PST_CONSISTENT:
				// UniqueToConsistent == AdditiveToConsistent on a vector whose slaves are zero
				if (has_storage_type(PST_UNIQUE) || has_storage_type(PST_ADDITIVE)) {
					UG_GPU_CHECK(ug4b200_additive_to_consistent(c, I, dev(), blockSize));
					set_storage_type(PST_CONSISTENT);
				} else return false;
				break;
			case PST_ADDITIVE:
				if (has_storage_type(PST_UNIQUE)) add_storage_type(PST_ADDITIVE);
				else if (has_storage_type(PST_CONSISTENT)) {
					UG_GPU_CHECK(ug4b200_set_slaves_zero(c, I, dev(), blockSize));
					set_storage_type(PST_ADDITIVE | PST_UNIQUE);
				} else return false;
				break;
			case PST_UNIQUE:
				if (has_storage_type(PST_ADDITIVE)) {
					UG_GPU_CHECK(ug4b200_additive_to_unique(c, I, dev(), blockSize));
					add_storage_type(PST_UNIQUE);
				} else if (has_storage_type(PST_CONSISTENT)) {
					UG_GPU_CHECK(ug4b200_set_slaves_zero(c, I, dev(), blockSize));
					set_storage_type(PST_ADDITIVE | PST_UNIQUE);
				} else return false;
				break;
			default: return false;
		}
		return true;
	}

	/// The smoothing kernel that is about to produce this (additive) vector may store its interface
	/// rows into the neighbours' peer windows itself; the change_storage_type(PST_CONSISTENT) that
	/// follows then only waits for the neighbours and adds the copies (ug4b200_interface_arm).
	void arm_consistent_push()
	{
		if (m_layouts) UG_GPU_CHECK(ug4b200_interface_arm(GPUManager::ctx(), m_layouts->iface(), dev()));
	}

	void swap(this_type& o)
	{
		std::swap(m_size, o.m_size); std::swap(m_dev, o.m_dev); std::swap(m_host, o.m_host);
		std::swap(m_hostValid, o.m_hostValid); std::swap(m_devValid, o.m_devValid);
		std::swap(m_layouts, o.m_layouts); std::swap(m_type, o.m_type);
	}

	/// one device double per process for reductions that must pass through an all-reduce
	static double* scalar_slot() { return GPUManager::scalar_slot(); }

  protected:
	static void add_block(double& a, const double& b) { a += b; }
	template <class X> static void add_block(X& a, const X& b) { for (size_t t = 0; t < X::size(); ++t) a[t] += b[t]; }
	double* dev_w() { m_devValid = true; m_hostValid = false; return m_dev; } // overwrite: no mirror needed
	void check_size(const this_type& v) const { UG_COND_THROW(v.m_size != m_size, "GPUVector: size mismatch " << m_size << " vs " << v.m_size); }
	void to_host()
	{
		if (m_hostValid) return;
		m_host.resize(m_size);
		if (m_size) UG_GPU_CHECK(ug4b200_d2h(GPUManager::ctx(), (void*)&m_host[0], m_dev, len() * sizeof(double)));
		m_hostValid = true;
	}
	void to_device()
	{
		if (m_devValid) return;
		if (m_size) {
			UG_GPU_CHECK(ug4b200_h2d(GPUManager::ctx(), m_dev, (const void*)&m_host[0], len() * sizeof(double)));
			UG_GPU_CHECK(ug4b200_sync(GPUManager::ctx()));
		}
		m_devValid = true;
	}
	void destroy()
	{
		if (m_dev) GPUManager::release(m_dev, len());
		m_dev = nullptr; m_size = 0; m_host.clear(); m_hostValid = false; m_devValid = true;
	}

	size_t m_size = 0;
	double* m_dev = nullptr;
	std::vector<value_type> m_host;
	bool m_hostValid = false, m_devValid = true;
	SmartPtr<GPUAlgebraLayouts> m_layouts;
	unsigned m_type = PST_UNDEFINED;
};

// ---- free functions (common/operations_vec.h:49-175; legacy overloads gpuvector.h:260-290) ----
template <typename T>
inline void VecScaleAssign(GPUVector<T>& dest, double alpha1, const GPUVector<T>& v1)
{ UG_GPU_CHECK(ug4b200_vec_scale_add2(GPUManager::ctx(), dest.len(), dest.dev(), alpha1, v1.dev(), 0.0, v1.dev())); }
template <typename T> inline void VecAssign(GPUVector<T>& dest, const GPUVector<T>& v1) { dest = v1; }
template <typename T>
inline void VecScaleAdd(GPUVector<T>& dest, double alpha1, const GPUVector<T>& v1, double alpha2, const GPUVector<T>& v2)
{
	UG_GPU_CHECK(ug4b200_vec_scale_add2(GPUManager::ctx(), dest.len(), dest.dev(), alpha1, v1.dev(), alpha2, v2.dev()));
	dest.set_storage_type(v1.get_storage_mask() & v2.get_storage_mask());
}
template <typename T>
inline void VecScaleAdd(GPUVector<T>& dest, double alpha1, const GPUVector<T>& v1, double alpha2, const GPUVector<T>& v2,
                        double alpha3, const GPUVector<T>& v3)
{
	UG_GPU_CHECK(ug4b200_vec_scale_add3(GPUManager::ctx(), dest.len(), dest.dev(), alpha1, v1.dev(), alpha2, v2.dev(), alpha3, v3.dev()));
	dest.set_storage_type(v1.get_storage_mask() & v2.get_storage_mask() & v3.get_storage_mask());
}
template <typename T> inline double VecProd(GPUVector<T>& a, GPUVector<T>& b) { return a.dotprod(b); }

} // namespace ug
