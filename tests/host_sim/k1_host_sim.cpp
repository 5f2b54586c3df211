// k1_host_sim.cpp -- TEST HARNESS ONLY (never part of libbrdae.so, never imported by the
// package).  It compiles the *device* source of the thread-per-system kernel
// (dae_scipy_b200/csrc/radau_k1.cuh) with the host compiler and runs the lane loop as a
// "warp of one lane", so that the kernel's control flow and arithmetic can be checked here,
// on a box without a GPU, bit for bit against the scalar C restatement in oracle/ --
// before GPU minutes are spent.  The product path has no CPU fallback: libbrdae.so launches
// CUDA kernels or returns an error.
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

// how often the kernel's fast paths and their fall-backs ran (tests assert that both sides are exercised):
// 0 optimistic LU attempts, 1 of them abandoned, 2 general LUs, 3 Newton passes on static addresses, 4 on pivoted ones,
// 5 first-iteration convergence tests
static long g_sim_counts[8];
extern "C" long *sim_counts() { return g_sim_counts; }
#define BRDAE_SIM_COUNT(i) (++g_sim_counts[i])
// The harness runs a warp of one lane, so the warp-wide votes of the kernel (every lane static? nobody pivoted?) only ever see
// that lane.  On the device a neighbouring lane can veto a fast path for the whole warp at any time; with sim_neighbour(p) set,
// an imaginary neighbour does so with probability p / 256 per vote (results must not depend on it).
static unsigned g_sim_neighbour_p = 0, g_sim_lcg = 12345u;
extern "C" void sim_neighbour(unsigned p, unsigned seed) { g_sim_neighbour_p = p; g_sim_lcg = seed; }
static inline bool sim_neighbour_agrees() {
    if (g_sim_neighbour_p == 0) return true;
    g_sim_lcg = g_sim_lcg * 1664525u + 1013904223u;
    return ((g_sim_lcg >> 16) & 255u) >= g_sim_neighbour_p;
}
#define BRDAE_SIM_NEIGHBOUR() sim_neighbour_agrees()
#include "../../dae_scipy_b200/csrc/host_args.hpp"
#include "../../dae_scipy_b200/csrc/radau_k1.cuh"
#ifdef BRDAE_USER_HEADER        // a user problem (dae_scipy_b200/plugin.py): the same generated functor header as the plugin build
#include BRDAE_USER_HEADER
namespace brdae {
const ProblemInfo *user_problem_info() {
    static const ProblemInfo pi{UserProblem::name(), BRDAE_USER, UserProblem::N, UserProblem::NP};
    return &pi;
}
const ProblemInfo *user_band_problem_info() { return nullptr; }
}
#else
namespace brdae { const ProblemInfo *user_problem_info() { return nullptr; } const ProblemInfo *user_band_problem_info() { return nullptr; } }
#endif

using namespace brdae;

template <class Prob, class Mass>
static void run(const SolveArgs &a) {
    // both layouts of the kernel (cold state next to the factors, or in its own global-memory slice) and
    // the finite-difference Jacobian mode
    const bool cold = std::getenv("BRDAE_SIM_COLD_GLOBAL") != nullptr;
    auto go = [&](auto cold_c, auto fd_c) {
        constexpr bool COLD = decltype(cold_c)::value, FD = decltype(fd_c)::value;
        using L = K1Layout<Prob, COLD, FD>;
        std::vector<double> hot(L::SHARED_SLOTS, 0.0), glob(L::GLOBAL_SLOTS + 1, 0.0);
        Slots<1> sm{hot.data()}, gj{glob.data()}, cd{COLD ? glob.data() : hot.data() + L::HOT};
        if (a.n_events > 0 || a.rep_cap > 0) k1_lane_loop<Prob, Mass, false, COLD, FD, true>(a, sm, cd, gj, a.t_eval);     // instrumented: events, reports
        else k1_lane_loop<Prob, Mass, false, COLD, FD, false>(a, sm, cd, gj, a.t_eval);
    };
    if (cold && a.jac_fd) go(std::true_type{}, std::true_type{});
    else if (cold) go(std::true_type{}, std::false_type{});
    else if (a.jac_fd) go(std::false_type{}, std::true_type{});
    else go(std::false_type{}, std::false_type{});
}

template <class Prob>
static void run_mass(bool dense, const SolveArgs &a) {
    if (dense) run<Prob, MassDense<Prob::N>>(a);
    else run<Prob, typename Prob::Mass>(a);
}

extern "C" int sim_solve(const brdae_batch *b, const brdae_options *o) {
    int rc = validate(b, o);
    if (rc != BRDAE_OK) return rc;
    if (b->n_systems == 0) return BRDAE_OK;
    unsigned long long ws[kCounterBytes / 8] = {0};
    SolveArgs a;
    fill_common(a, b, o, ws);
    const bool dense = b->mass != nullptr;
    switch (b->problem) {
#ifdef BRDAE_USER_HEADER            // the harness of a user problem holds that problem only (compiles in seconds)
    case BRDAE_USER: run_mass<UserProblem>(dense, a); break;
#else
    case BRDAE_PENDULUM3: run_mass<Pendulum<3>>(dense, a); break;
    case BRDAE_PENDULUM2: run_mass<Pendulum<2>>(dense, a); break;
    case BRDAE_PENDULUM1: run_mass<Pendulum<1>>(dense, a); break;
    case BRDAE_ROBERTSON: run_mass<Robertson>(dense, a); break;
    case BRDAE_TRANSISTOR: run_mass<Transistor>(dense, a); break;
    case BRDAE_ROBERTSON_ODE: run_mass<RobertsonODE>(dense, a); break;
    case BRDAE_JAY: run_mass<Jay>(dense, a); break;
    case BRDAE_POLYDAE2: run_mass<PolyDAE2>(dense, a); break;
    case BRDAE_LINEAR7: run_mass<Linear7>(dense, a); break;
#endif
    default: return BRDAE_ERR_UNSUPPORTED;
    }
    return BRDAE_OK;
}
