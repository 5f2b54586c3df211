// k1_transistor.cu -- K1 instantiations for the transistor amplifier (n = 5, singular
// non-diagonal mass matrix built from each system's capacitances).
#include "k1_launch.cuh"

// launch geometry: one 256-thread CTA per SM (two 128-thread CTAs: 140.5 against 135.5 ms; four 64-thread CTAs: 196.6 ms --
// GPU call 18 of round 2)
#ifndef K1_TRA_BLOCK
#define K1_TRA_BLOCK 256
#endif
#ifndef K1_TRA_MINB
#define K1_TRA_MINB 1
#endif

namespace brdae {

int k1_transistor(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    if (dense_mass) return k1_launch<Transistor, MassDense<5>, K1_TRA_BLOCK, K1_TRA_MINB, true>(a, sms, s, g, geom_only);
    return k1_launch<Transistor, Transistor::Mass, K1_TRA_BLOCK, K1_TRA_MINB, true>(a, sms, s, g, geom_only);
}

}  // namespace brdae
