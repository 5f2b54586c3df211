// k1_launch.cuh -- launch helper shared by the k1_*.cu translation units.
#pragma once
#include <cstdlib>

#include "kernels.cuh"
#include "radau_k1.cuh"

namespace brdae {

template <class Prob, class Mass, int BLOCK, int MIN_BLOCKS, bool CTA_WIDE, bool COLD_GLOBAL, bool FD, bool EVENTS>
int k1_launch_mode(const SolveArgs &a_in, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only, int newton_reps) {
    // Newton passes per cycle: a scheduling knob tuned per problem family (results do not depend on it);
    // BRDAE_NEWTON_REPS overrides it for experiments
    SolveArgs a = a_in;
    a.newton_reps = newton_reps;
    if (const char *e = std::getenv("BRDAE_NEWTON_REPS")) { const int v = std::atoi(e); if (v >= 1 && v <= 16) a.newton_reps = v; }
    using L = K1Layout<Prob, COLD_GLOBAL, FD>;
    constexpr int smem = L::SHARED_SLOTS * BLOCK * (int)sizeof(double);
    int64_t want = (a.B + BLOCK - 1) / BLOCK;
    int64_t cap = (int64_t)(sms > 0 ? sms : 148) * MIN_BLOCKS;  // persistent grid: resident CTAs only
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    if (g) { g->grid = grid; g->block = BLOCK; g->smem = smem; g->scratch_bytes = (size_t)grid * BLOCK * L::GLOBAL_SLOTS * sizeof(double); }
    if (geom_only) return BRDAE_OK;
    auto kern = k1_kernel<Prob, Mass, BLOCK, MIN_BLOCKS, CTA_WIDE, COLD_GLOBAL, FD, EVENTS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return BRDAE_ERR_CUDA;
    kern<<<grid, BLOCK, smem, s>>>(a);
    return cudaPeekAtLastError() == cudaSuccess ? BRDAE_OK : BRDAE_ERR_CUDA;
}

// analytic / constant Jacobian, or the finite-difference instantiation when the batch asks for it
template <class Prob, class Mass, int BLOCK, int MIN_BLOCKS, bool CTA_WIDE = false, bool COLD_GLOBAL = false>
int k1_launch(const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only, int newton_reps = 3) {
    // device-side events (brdae_batch.n_events > 0) are their own instantiations: the event code costs the plain kernels nothing
    // ... and so is the per-attempt report (brdae_batch.report_cap > 0): the same instrumented instantiation writes it
    if (a.n_events > 0 || a.rep_cap > 0) {
        if (a.jac_fd) return k1_launch_mode<Prob, Mass, BLOCK, MIN_BLOCKS, CTA_WIDE, COLD_GLOBAL, true, true>(a, sms, s, g, geom_only, newton_reps);
        return k1_launch_mode<Prob, Mass, BLOCK, MIN_BLOCKS, CTA_WIDE, COLD_GLOBAL, false, true>(a, sms, s, g, geom_only, newton_reps);
    }
    if (a.jac_fd) return k1_launch_mode<Prob, Mass, BLOCK, MIN_BLOCKS, CTA_WIDE, COLD_GLOBAL, true, false>(a, sms, s, g, geom_only, newton_reps);
    return k1_launch_mode<Prob, Mass, BLOCK, MIN_BLOCKS, CTA_WIDE, COLD_GLOBAL, false, false>(a, sms, s, g, geom_only, newton_reps);
}

}  // namespace brdae
