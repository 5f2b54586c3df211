// k2_user.cu -- the banded CTA-per-system kernel K2 (k2_band.cuh) for a USER band problem with more than 8 unknowns (plugin builds).
//
// The reference integrates any `fun` (radauDAE.py:263-266), by default with a finite-difference Jacobian and, for large
// systems, a banded `jac_sparsity` (radauDAE.py:468-481; recursive_pendulum.py:165-166).  A mid-size banded system -- a
// semi-discretised PDE, a chain -- is stated row by row:
//
//     struct UserBand {                          // in the header named by -DBRDAE_USER_BAND_HEADER="\"...\""
//         static constexpr int N, NP, KL, KU;    // unknowns, parameters per system, lower / upper bandwidth of df/dy
//         static const char *name();
//         static double rhs_row(int i, double t, const double *y, const double *p, int n);   // f_i(t, y)
//         static bool mass(int i, int n);        // M = diag(0/1): is row i differential?
//     };
//
// One CTA integrates one system (rows spread over the threads); the Jacobian is scipy's num_jac with column groups on the band
// (kl + ku + 1 right-hand-side evaluations), the iteration matrices are factorised by the pivoted band LU.  dae_scipy_b200/plugin.py
// (`compile_band_problem`) generates the header and links this unit, compiled with it, in place of the stub below.
#include "host_args.hpp"
#include "kernels.cuh"

#ifdef BRDAE_USER_BAND_HEADER
#include BRDAE_USER_BAND_HEADER

#define K2_KL UserBand::KL
#define K2_KU UserBand::KU
#define K2_CHAIN 0
#define K2_ENTRY k2_user
#define K2_SCRATCH_FN k2_user_scratch_bytes
#include "k2_band.cuh"

namespace brdae {
const ProblemInfo *user_band_problem_info() {
    static const ProblemInfo pi{UserBand::name(), BRDAE_USER, UserBand::N, UserBand::NP};
    return &pi;
}
}  // namespace brdae

#else  // stock library: no user band problem

namespace brdae {
int k2_user(bool, const SolveArgs &, const brdae_batch *, int, cudaStream_t, LaunchGeom *, bool) { return BRDAE_ERR_UNKNOWN_PROBLEM; }
size_t k2_user_scratch_bytes(int, int, int64_t, bool) { return 0; }
const ProblemInfo *user_band_problem_info() { return nullptr; }
}  // namespace brdae

#endif
