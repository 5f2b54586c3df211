// k1_pendulum.cu -- K1 instantiations for the single pendulum, index 3 / 2 / 1.
#include "k1_launch.cuh"

// launch geometry (overridable for tuning experiments): two 128-thread CTAs per SM.  With the fixed cycle a
// CTA synchronises once per cycle, and two half-size CTAs wait less at that barrier than one 256-thread CTA
// (1M systems: 71.9 ms against 73.9 ms; four 64-thread CTAs lose the shared instruction stream: 96.6 ms).
#ifndef K1_PEND_BLOCK
#define K1_PEND_BLOCK 128
#endif
#ifndef K1_PEND_MINB
#define K1_PEND_MINB 2
#endif
#ifndef K1_PEND_CTAWIDE
#define K1_PEND_CTAWIDE true
#endif
#ifndef K1_PEND_COLD
#define K1_PEND_COLD false
#endif

namespace brdae {

template <int INDEX>
static int go(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    using Prob = Pendulum<INDEX>;
    if (dense_mass) return k1_launch<Prob, MassDense<5>, K1_PEND_BLOCK, K1_PEND_MINB, K1_PEND_CTAWIDE, K1_PEND_COLD>(a, sms, s, g, geom_only, 4);
    return k1_launch<Prob, typename Prob::Mass, K1_PEND_BLOCK, K1_PEND_MINB, K1_PEND_CTAWIDE, K1_PEND_COLD>(a, sms, s, g, geom_only, 4);
}

int k1_pendulum(int index, bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    if (index == 3) return go<3>(dense_mass, a, sms, s, g, geom_only);
    if (index == 2) return go<2>(dense_mass, a, sms, s, g, geom_only);
    if (index == 1) return go<1>(dense_mass, a, sms, s, g, geom_only);
    return BRDAE_ERR_UNKNOWN_PROBLEM;
}

}  // namespace brdae
