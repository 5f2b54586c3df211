// k1_host_sim.cpp -- TEST HARNESS ONLY (never part of libbrdae.so, never imported by the
// package).  It compiles the *device* source of the thread-per-system kernel
// (dae_scipy_b200/csrc/radau_k1.cuh) with the host compiler and runs the lane loop as a
// "warp of one lane", so that the kernel's control flow and arithmetic can be checked here,
// on a box without a GPU, bit for bit against the scalar C restatement in oracle/ --
// before GPU minutes are spent.  The product path has no CPU fallback: libbrdae.so launches
// CUDA kernels or returns an error.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../dae_scipy_b200/csrc/host_args.hpp"
#include "../../dae_scipy_b200/csrc/radau_k1.cuh"

using namespace brdae;

template <class Prob, class Mass>
static void run(const SolveArgs &a) {
    // both layouts of the kernel: cold state next to the factors, or in its own (global-memory) slice
    if (!std::getenv("BRDAE_SIM_COLD_GLOBAL")) {
        using L = K1Layout<Prob, false>;
        std::vector<double> slots(L::SHARED_SLOTS, 0.0);
        Slots<1> sm{slots.data()}, cd{slots.data() + L::HOT};
        k1_lane_loop<Prob, Mass, false, false>(a, sm, cd, a.t_eval);
    } else {
        using L = K1Layout<Prob, true>;
        std::vector<double> hot(L::SHARED_SLOTS, 0.0), cold(L::COLD, 0.0);
        Slots<1> sm{hot.data()}, cd{cold.data()};
        k1_lane_loop<Prob, Mass, false, true>(a, sm, cd, a.t_eval);
    }
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
    case BRDAE_PENDULUM3: run_mass<Pendulum<3>>(dense, a); break;
    case BRDAE_PENDULUM2: run_mass<Pendulum<2>>(dense, a); break;
    case BRDAE_PENDULUM1: run_mass<Pendulum<1>>(dense, a); break;
    case BRDAE_ROBERTSON: run_mass<Robertson>(dense, a); break;
    case BRDAE_TRANSISTOR: run_mass<Transistor>(dense, a); break;
    case BRDAE_ROBERTSON_ODE: run_mass<RobertsonODE>(dense, a); break;
    case BRDAE_JAY: run_mass<Jay>(dense, a); break;
    case BRDAE_POLYDAE2: run_mass<PolyDAE2>(dense, a); break;
    case BRDAE_LINEAR7: run_mass<Linear7>(dense, a); break;
    default: return BRDAE_ERR_UNSUPPORTED;
    }
    return BRDAE_OK;
}
