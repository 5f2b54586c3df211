// k1_user.cu -- K1 instantiation for a USER-SUPPLIED problem functor (plugin builds).
//
// The reference takes the right-hand side and the Jacobian as Python closures (`fun`, `jac` of
// solve_ivp / RadauDAE.__init__, radauDAE.py:263-266, 464-513).  On the device they are a functor
//
//     struct UserProblem {                       // in the header named by -DBRDAE_USER_HEADER="\"...\""
//         static constexpr int N, NP, BLOCK;     // unknowns (<= 8), parameters per system, threads per CTA
//         static const char *name();
//         struct P { ... };                      // one system's parameters
//         using Mass = MassDiag01<N, ONES>;      // its own mass matrix diag(0/1) (dense ones: mass_matrix=)
//         static void load(P &, const double *params, int64_t stride, int64_t offset);
//         static void rhs(const P &, double t, const double (&y)[N], double (&f)[N]);
//         static void jac(const P &, double t, const double (&y)[N], double (&J)[N][N]);
//     };
//
// and a plugin build links this translation unit, compiled WITH the header, in place of the stub below:
// the result is a complete libbrdae (same C ABI, include/brdae.h) whose problem BRDAE_USER is that functor,
// with every mode of the thread-per-system kernel (analytic / constant / finite-difference Jacobian, own or
// dense mass matrix, t_eval, dense output).  dae_scipy_b200/plugin.py generates the header from a few lines
// of CUDA C++ and runs nvcc; tests/host_sim compiles the same header for the host.
#include "host_args.hpp"
#include "k1_launch.cuh"

#ifdef BRDAE_USER_HEADER
#include BRDAE_USER_HEADER

namespace brdae {

static_assert(UserProblem::N >= 1 && UserProblem::N <= BRDAE_NMAX_SMALL, "thread-per-system kernels: 1 <= N <= 8");

int k1_user(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    if (dense_mass) return k1_launch<UserProblem, MassDense<UserProblem::N>, UserProblem::BLOCK, 1, true>(a, sms, s, g, geom_only);
    return k1_launch<UserProblem, UserProblem::Mass, UserProblem::BLOCK, 1, true>(a, sms, s, g, geom_only);
}

const ProblemInfo *user_problem_info() {
    static const ProblemInfo pi{UserProblem::name(), BRDAE_USER, UserProblem::N, UserProblem::NP};
    return &pi;
}

}  // namespace brdae

#else  // stock library: no user problem

namespace brdae {

int k1_user(bool, const SolveArgs &, int, cudaStream_t, LaunchGeom *, bool) { return BRDAE_ERR_UNKNOWN_PROBLEM; }
const ProblemInfo *user_problem_info() { return nullptr; }

}  // namespace brdae

#endif
