/*
 * oracle/ref_prelude.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Force-included (-include) into every translation unit of oracle/_ref so that two more pieces of the
 * REAL reference compile without boost:
 *   lib_algebra/operator/preconditioner/ilu.h        FactorizeILUSorted, FactorizeILUBeta, invert_L, invert_U
 *   lib_algebra/ordering_strategies/algorithms/native_cuthill_mckee.cpp   ComputeCuthillMcKeeOrder
 * Both reach boost::graph only through native_cuthill_mckee.h -> util.h (adjacency_list adapters that
 * neither of them uses).  Defining that header's include guard skips it; the two names the code really
 * needs from it are declared here (declarations only — no reference source is copied).
 */
#ifndef ORACLE_REF_PRELUDE_H
#define ORACLE_REF_PRELUDE_H
#ifdef __cplusplus
#include <cstddef>
#include <vector>
#define UG_BASE_LIB_ALGEBRA_ORDERING_STRATEGIES_ALGORITHMS_NATIVE_CUTHILL_MCKEE_H
namespace ug {
template <typename TAlgebra, typename O_t> class NativeCuthillMcKeeOrdering;   // named inside ILU::set_sort only
void ComputeCuthillMcKeeOrder(std::vector<size_t>& vNewIndex, std::vector<std::vector<size_t> >& vvNeighbour,
                              bool bReverse, bool bPreserveConsec);             // native_cuthill_mckee.h:88-91
}
#endif
#endif
