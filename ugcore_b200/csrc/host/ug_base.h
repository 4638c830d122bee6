// ug_base.h — minimal stand-ins for the ugcore infrastructure the GPU algebra touches
// (UG_THROW, SmartPtr, small-block types) plus the per-process device manager.
//
// In a real ugcore build these come from ugbase/common (error.h:57-85, smart_pointer.h)
// and ugbase/lib_algebra/small_algebra; names and meaning are kept so the classes in
// this directory read like their CPU counterparts and can be moved into
// ugbase/lib_algebra/gpu_algebra/ unchanged.
#pragma once
#ifdef UG4B200_WITH_UGCORE
#include "ug4b200.h"                       // on the include path of the ugcore build (integration/0005-cuda_cmake.patch)
#else
#include "../../../include/ug4b200.h"
#endif
#include <cstddef>
#include <cstdlib>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef UG4B200_WITH_UGCORE
// ---- inside a ugcore build (-I<ug4>/ugcore/ugbase): the real infrastructure, no stand-ins ----
//   UG_THROW / UG_COND_THROW / THROW_IF_NOT_EQUAL   ugbase/common/error.h:57-85, 61, 181
//   SmartPtr / ConstSmartPtr / make_sp(T*)          ugbase/common/util/smart_pointer.h:109-260, 839-850
//   DenseMatrix / DenseVector / FixedArray1/2       ugbase/lib_algebra/small_algebra/small_algebra.h
//   AlgebraType (has GPU = 1)                       ugbase/lib_algebra/algebra_type.h:50-60
//   ParallelStorageType                             ugbase/lib_algebra/parallelization/parallel_storage_type.h:65-71
// tests/boundary/ compiles this configuration against /root/reference (tests/test_boundary.py).
#include "common/common.h"
#include "common/util/smart_pointer.h"
#include "lib_algebra/small_algebra/small_algebra.h"
#include "lib_algebra/algebra_type.h"
#include "lib_algebra/parallelization/parallel_storage_type.h"

// ugcore's SmartPtr and make_sp(T*) live in the GLOBAL namespace (smart_pointer.h:109, 839-850); the helpers go next to
// them — a ug::make_sp would hide ugcore's own inside namespace ug (linear_iterator.h:169 calls make_sp(new ...))
/// make_sp<T>(ctor args...): this overload does the `new`
template <class T, class... Args> SmartPtr<T> make_sp(Args&&... a) { return SmartPtr<T>(new T(std::forward<Args>(a)...)); }
/// SmartPtr::cast_dynamic under a name that also exists for the stand-alone build's std::shared_ptr
template <class TDest, class T> SmartPtr<TDest> sp_cast_dynamic(const SmartPtr<T>& p) { return p.template cast_dynamic<TDest>(); }

#else
// ---- stand-alone build: minimal stand-ins with ugcore's names and meaning ----
namespace ug {

typedef double number;

/// ugbase/common/error.h: UGError carries a message stack; one message suffices here
class UGError : public std::runtime_error {
  public:
	explicit UGError(const std::string& m) : std::runtime_error(m) {}
};

#define UG_THROW(msg)                                                    \
	do { std::stringstream ss_; ss_ << msg; throw ::ug::UGError(ss_.str()); } while (0)
#define UG_COND_THROW(cond, msg) do { if (cond) UG_THROW(msg); } while (0)
#define THROW_IF_NOT_EQUAL(a, b) UG_COND_THROW((a) != (b), #a << " != " << #b << " (" << (a) << " vs " << (b) << ")")

template <class T> using SmartPtr = std::shared_ptr<T>;
template <class T> using ConstSmartPtr = std::shared_ptr<const T>;
template <class T, class... Args> SmartPtr<T> make_sp(Args&&... a) { return std::make_shared<T>(std::forward<Args>(a)...); }
template <class TDest, class T> SmartPtr<TDest> sp_cast_dynamic(const SmartPtr<T>& p) { return std::dynamic_pointer_cast<TDest>(p); }

// ---- small algebra (ugbase/lib_algebra/small_algebra): column-major fixed blocks ----
template <class T, size_t N> struct FixedArray1 {
	T values[N];
	T& operator[](size_t i) { return values[i]; }
	const T& operator[](size_t i) const { return values[i]; }
	static size_t size() { return N; }
};
template <class T, size_t R, size_t C> struct FixedArray2 { // ColMajor, fixed_array_impl.h:182-203
	T values[R * C];
	T& operator()(size_t r, size_t c) { return values[r + R * c]; }
	const T& operator()(size_t r, size_t c) const { return values[r + R * c]; }
	static size_t num_rows() { return R; }
	static size_t num_cols() { return C; }
};
template <class A> struct DenseVector : public A {
	DenseVector& operator=(double d) { for (size_t i = 0; i < A::size(); ++i) (*this)[i] = d; return *this; }
};
template <class A> struct DenseMatrix : public A {
	DenseMatrix& operator=(double d)
	{
		for (size_t r = 0; r < A::num_rows(); ++r) for (size_t c = 0; c < A::num_cols(); ++c) (*this)(r, c) = (r == c) ? d : 0.0;
		return *this;
	}
};

template <class T> struct block_traits;
template <> struct block_traits<double> { enum { static_size = 1, static_num_rows = 1 }; };
template <size_t N> struct block_traits<DenseVector<FixedArray1<double, N> > > { enum { static_size = N, static_num_rows = N }; };
template <size_t N> struct block_traits<DenseMatrix<FixedArray2<double, N, N> > > { enum { static_size = N * N, static_num_rows = N }; };

// ---- AlgebraType (ugbase/lib_algebra/algebra_type.h:53-57) ----
struct AlgebraType {
	enum Type { CPU = 0, GPU = 1 };
	AlgebraType(Type t, int bs) : m_type(t), m_blockSize(bs) {}
	int type() const { return m_type; }
	int blocksize() const { return m_blockSize; }
	int m_type, m_blockSize;
};

// ---- parallel storage types (lib_algebra/parallelization/parallel_storage_type.h:65-71) ----
enum ParallelStorageType { PST_UNDEFINED = 0, PST_CONSISTENT = 1, PST_ADDITIVE = 2, PST_UNIQUE = 4 };

} // namespace ug
#endif // UG4B200_WITH_UGCORE

// ---- profiler zones (ugbase/common/profiler/profiler.h: PROFILE_BEGIN_GROUP / PROFILE_FUNC_GROUP; the GMG's own
// GMG_PROFILE_BEGIN(GMG_PreSmooth) ..., mg_solver_impl.hpp:60-69) — as NVTX ranges with ugcore's zone names, so that an
// Nsight Systems timeline of a solve reads like ugcore's profiler output.  nvtx3 is header-only and resolves its
// injection library at run time (no link dependency); without the header the zones compile to nothing.
#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>) && !defined(UG4B200_NO_NVTX)
#include <nvtx3/nvToolsExt.h>
#define UG4B200_HAVE_NVTX 1
#endif
#endif
namespace ug {
struct GPUProfileZone {
#ifdef UG4B200_HAVE_NVTX
	explicit GPUProfileZone(const char* name) { nvtxRangePushA(name); }
	~GPUProfileZone() { nvtxRangePop(); }
#else
	explicit GPUProfileZone(const char*) {}
#endif
	GPUProfileZone(const GPUProfileZone&) = delete;
	GPUProfileZone& operator=(const GPUProfileZone&) = delete;
};
} // namespace ug
#define UG_GPU_ZONE_CAT2(a, b) a##b
#define UG_GPU_ZONE_CAT(a, b) UG_GPU_ZONE_CAT2(a, b)
/// a zone that lasts until the end of the enclosing scope, named like ugcore's (GMG_PROFILE_BEGIN(GMG_PreSmooth))
#define UG_GPU_ZONE(name) ::ug::GPUProfileZone UG_GPU_ZONE_CAT(ugGpuZone_, __LINE__)(#name)

namespace ug {

/// Turns a C-ABI error code into a UGError (what CUDA_CHECK_STATUS did, cuda_manager.h:61-77)
#define UG_GPU_CHECK(call)                                                                  \
	do {                                                                                    \
		int rc_ = (call);                                                                   \
		if (rc_ != 0) UG_THROW(#call << " failed (" << rc_ << "): " << ug4b200_last_error(::ug::GPUManager::ctx_or_null())); \
	} while (0)

/// One device context per process / MPI rank (replaces CUDAManager, cuda_manager.cpp:54-164).
class GPUManager {
  public:
	/// device < 0: LOCAL_RANK % device count (one rank per GPU)
	static void init(int device = -1, void* stream = nullptr)
	{
		GPUManager& m = inst();
		if (m.m_ctx) return;
		if (device < 0) {
			const char* lr = std::getenv("LOCAL_RANK");
			device = lr ? std::atoi(lr) : 0;
		}
		ug4b200_ctx* c = nullptr;
		int rc = ug4b200_ctx_create(device, stream, &c);
		if (rc != 0) UG_THROW("GPUManager: cannot create device context: " << ug4b200_last_error(nullptr));
		m.m_ctx = c;
	}
	static ug4b200_ctx* ctx() { if (!inst().m_ctx) init(); return inst().m_ctx; }
	static ug4b200_ctx* ctx_or_null() { return inst().m_ctx; }
	static void finalize()
	{
		GPUManager& m = inst();
		m.release_pool();
		if (m.m_ctx && m.m_scalarSlot) ug4b200_free(m.m_ctx, m.m_scalarSlot);
		m.m_scalarSlot = nullptr;
		if (m.m_ctx) { ug4b200_ctx_destroy(m.m_ctx); m.m_ctx = nullptr; }
		bump_generation();
	}
	/// pooled device allocation: solver work vectors are cloned per apply (cg.h:120-122)
	static double* alloc(size_t n)
	{
		GPUManager& m = inst();
		auto it = m.m_pool.find(n);
		if (it != m.m_pool.end() && !it->second.empty()) { double* p = it->second.back(); it->second.pop_back(); return p; }
		void* p = nullptr;
		int rc = ug4b200_alloc(ctx(), n * sizeof(double), &p);
		if (rc != 0) UG_THROW("GPUManager: device allocation of " << n * sizeof(double) << " bytes failed: " << ug4b200_last_error(ctx()));
		return (double*)p;
	}
	/// back into the pool — unless the context is gone (a vector that outlives finalize()): its memory went with the context
	static void release(double* p, size_t n) { if (p && inst().m_ctx) inst().m_pool[n].push_back(p); }
	/// a few device doubles per process for reductions that pass through an all-reduce; owned here so that
	/// finalize() / a new context never leave a dangling pointer behind
	static double* scalar_slot()
	{
		GPUManager& m = inst();
		if (!m.m_scalarSlot) m.m_scalarSlot = (double*)alloc_bytes(8 * sizeof(double));
		return m.m_scalarSlot;
	}
	static void* alloc_bytes(size_t bytes)
	{
		void* p = nullptr;
		int rc = ug4b200_alloc(ctx(), bytes, &p);
		if (rc != 0) UG_THROW("GPUManager: device allocation failed: " << ug4b200_last_error(ctx()));
		return p;
	}
	static void free_bytes(void* p) { if (p && inst().m_ctx) ug4b200_free(inst().m_ctx, p); }
	/// pcl::ProcRank() / pcl::NumProcs() of this process (0 / 1 until the communicator is set up)
	static int proc_rank() { return inst().m_rank; }
	static int num_procs() { return inst().m_nranks; }
	static void set_procs(int nranks, int rank) { inst().m_nranks = nranks; inst().m_rank = rank; }
	/// Generation of the device-side solver data.  Bumped whenever a matrix mirror, a preconditioner's device
	/// buffers or a level hierarchy is (re)built or dropped; captured CUDA graphs bake those pointers in and are
	/// only replayed while the generation they were captured under is still current (solvers.h).
	static unsigned long long generation() { return inst().m_generation; }
	static void bump_generation() { ++inst().m_generation; }

  private:
	static GPUManager& inst() { static GPUManager m; return m; }
	void release_pool()
	{
		if (m_ctx) for (auto& kv : m_pool) for (double* p : kv.second) ug4b200_free(m_ctx, p);
		m_pool.clear();
	}
	ug4b200_ctx* m_ctx = nullptr;
	int m_rank = 0, m_nranks = 1;
	unsigned long long m_generation = 1;
	double* m_scalarSlot = nullptr;
	std::map<size_t, std::vector<double*> > m_pool;
};

} // namespace ug
