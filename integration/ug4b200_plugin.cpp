// integration/ug4b200_plugin.cpp — alternative delivery as a ugcore plugin (not compiled in this repository: the bridge
// headers need boost::mpl, which is not part of /root/reference).
//
// ugcore loads lib*.so from its plugin directory and calls
//     extern "C" void InitUGPlugin_<name>(ug::bridge::Registry*, std::string)
// (ugbase/common/util/plugin_util_dynamic.cpp:60-168; the symbol name is built at :136-140 from the library name,
// so this file must be linked into libug4b200.so's sibling `libug4b200plugin.so` -> InitUGPlugin_ug4b200plugin, or the
// plugin is named `ug4b200`).  The registration below is what bridge/algebra_bridges/*.cpp and
// bridge/disc_bridges/multigrid_bridge.cpp do for CompileAlgebraList, restricted to the GPU algebras — the route for
// an installation that cannot rebuild ugcore with -DGPU_ALGEBRA=ON (integration/0001..0007 patches).
#include "bridge/bridge.h"
#include "bridge/util.h"
#include "bridge/util_algebra_dependent.h"          // RegisterAlgebraDependent (:189-214)
#include "bridge/util_domain_algebra_dependent.h"   // RegisterDomainAlgebraDependent
#include "lib_algebra/gpu_algebra/ug4b200/multigrid.h"   // GPUAlgebra, GPUBlockAlgebra<N>, solvers, smoothers, GMG

namespace ug {
namespace ug4b200 {

/// the classes util.solver / solver_util.lua instantiate on this path (scripts/util/solver_util.lua:602-904), registered
/// under ugcore's own group names with the algebra suffix "GPU1" / "GPU3" (bridge/suffix_tag.h:88-140), so that
/// InitUG(dim, AlgebraType("GPU", 1)) resolves CG, BiCGStab, LinearSolver, GMRES, LU, Jacobi, GaussSeidel, ILU,
/// StdConvergenceCheck — exactly the names the Lua scripts use (solver_bridge.cpp:205-234, preconditioner_bridge.cpp)
struct Functionality {
	template <typename TAlgebra>
	static void Algebra(bridge::Registry& reg, std::string grp)
	{
		typedef typename TAlgebra::vector_type vector_type;
		const std::string suffix = bridge::GetAlgebraSuffix<TAlgebra>();
		const std::string tag = bridge::GetAlgebraTag<TAlgebra>();
		{
			typedef Jacobi<TAlgebra> T;
			typedef IPreconditioner<TAlgebra> TBase;
			std::string name = std::string("Jacobi").append(suffix);
			reg.add_class_<T, TBase>(name, grp, "Jacobi preconditioner (B200)")
			    .add_constructor()
			    .template add_constructor<void (*)(number)>("DampingFactor")
			    .add_method("set_block", &T::set_block, "", "block")
			    .set_construct_as_smart_pointer(true);
			reg.add_class_to_group(name, "Jacobi", tag);
		}
		{
			typedef GaussSeidel<TAlgebra> T;
			typedef IPreconditioner<TAlgebra> TBase;
			std::string name = std::string("GaussSeidel").append(suffix);
			reg.add_class_<T, TBase>(name, grp, "multicolour Gauss-Seidel (B200)")
			    .add_constructor()
			    .add_method("set_sor_relax", &T::set_sor_relax, "", "sor relaxation")
			    .set_construct_as_smart_pointer(true);
			reg.add_class_to_group(name, "GaussSeidel", tag);
		}
		{
			typedef ILU<TAlgebra> T;
			typedef IPreconditioner<TAlgebra> TBase;
			std::string name = std::string("ILU").append(suffix);
			reg.add_class_<T, TBase>(name, grp, "ILU(0) / ILU(beta) (B200)")
			    .add_constructor()
			    .add_method("set_beta", &T::set_beta, "", "beta")
			    .add_method("set_sort", &T::set_sort, "", "bSort")
			    .add_method("set_inversion_eps", &T::set_inversion_eps, "", "eps")
			    .set_construct_as_smart_pointer(true);
			reg.add_class_to_group(name, "ILU", tag);
		}
		{
			typedef CG<vector_type> T;
			typedef IPreconditionedLinearOperatorInverse<vector_type> TBase;
			std::string name = std::string("CG").append(suffix);
			reg.add_class_<T, TBase>(name, grp, "Conjugate Gradient (B200, device-resident)")
			    .add_constructor()
			    .set_construct_as_smart_pointer(true);
			reg.add_class_to_group(name, "CG", tag);
		}
		{
			typedef BiCGStab<vector_type> T;
			typedef IPreconditionedLinearOperatorInverse<vector_type> TBase;
			std::string name = std::string("BiCGStab").append(suffix);
			reg.add_class_<T, TBase>(name, grp, "BiCGStab (B200)")
			    .add_constructor()
			    .add_method("set_restart", &T::set_restart)
			    .add_method("set_min_orthogonality", &T::set_min_orthogonality)
			    .set_construct_as_smart_pointer(true);
			reg.add_class_to_group(name, "BiCGStab", tag);
		}
		{
			typedef LinearSolver<vector_type> T;
			typedef IPreconditionedLinearOperatorInverse<vector_type> TBase;
			std::string name = std::string("LinearSolver").append(suffix);
			reg.add_class_<T, TBase>(name, grp, "Linear Solver (B200)")
			    .add_constructor()
			    .set_construct_as_smart_pointer(true);
			reg.add_class_to_group(name, "LinearSolver", tag);
		}
		{
			typedef LU<TAlgebra> T;
			typedef ILinearOperatorInverse<vector_type> TBase;
			std::string name = std::string("LU").append(suffix);
			reg.add_class_<T, TBase>(name, grp, "dense LU base solver (host factorisation, device solve)")
			    .add_constructor()
			    .set_construct_as_smart_pointer(true);
			reg.add_class_to_group(name, "LU", tag);
		}
	}
};

} // namespace ug4b200
} // namespace ug

extern "C" void InitUGPlugin_ug4b200(ug::bridge::Registry* reg, std::string grp)
{
	grp.append("ug4b200/");
	typedef boost::mpl::list<ug::GPUAlgebra, ug::GPUBlockAlgebra<3>, ug::bridge::end_boost_list> GPUAlgebraList;
	try {
		ug::bridge::RegisterAlgebraDependent<ug::ug4b200::Functionality, GPUAlgebraList>(*reg, grp);
	}
	UG_REGISTRY_CATCH_THROW(grp);
}

/// ugbase/common/util/plugin_util_dynamic.cpp:171-187: called before the library is unloaded
extern "C" void FinalizeUGPlugin_ug4b200() { ug::GPUManager::finalize(); }
