// k4_host_sim.cpp -- TEST HARNESS ONLY (never part of libbrdae.so, never imported by the package).  It compiles the
// *device* source of the lane-per-rope chain kernel (dae_scipy_b200/csrc/chain_k4.cuh) with the host compiler and runs
// the lane loop as a "warp of one lane" (slot stride 1), so that the kernel's control flow and arithmetic are checked
// here, on a box without a GPU, bit for bit against the scalar C restatement in oracle/ -- before GPU minutes are
// spent.  The product path has no CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../dae_scipy_b200/csrc/host_args.hpp"
#include "../../dae_scipy_b200/csrc/chain_k4.cuh"

namespace brdae { const ProblemInfo *user_problem_info() { return nullptr; } const ProblemInfo *user_band_problem_info() { return nullptr; } }
using namespace brdae;

extern "C" int sim_solve_chain(const brdae_batch *b, const brdae_options *o) {
    int rc = validate(b, o);
    if (rc != BRDAE_OK) return rc;
    if (b->problem != BRDAE_NPENDULUM || b->mass || b->jac_const) return BRDAE_ERR_UNSUPPORTED;
    if (b->n_systems == 0) return BRDAE_OK;
    unsigned long long ws[kCounterBytes / 8] = {0};
    SolveArgs a;
    fill_common(a, b, o, ws);
    std::vector<int32_t> vidx(b->n, 0);
    if (b->var_index) for (int c = 0; c < b->n; ++c) vidx[c] = b->var_index[c];
    a.atol_dev = b->atol; a.var_index_dev = vidx.data();
    if (const char *e = std::getenv("BRDAE_NEWTON_REPS")) { const int v = std::atoi(e); if (v >= 1 && v <= 16) a.newton_reps = v; }
    std::vector<double> state((size_t)(b->n / 5) * k4::SPN, 0.0);
    k4::Rope<1> R{state.data()};
    k4::Stager<1, false> S;                 // the harness reads the nodes in place (staging is a device-side data movement)
    k4::k4_lane_loop<1>(a, R, true, S);        // the one lane pulls ropes from the work counter until none is left
    return BRDAE_OK;
}
