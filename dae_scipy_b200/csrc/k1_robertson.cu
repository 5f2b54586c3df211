// k1_robertson.cu -- K1 instantiations for the Robertson DAE (n = 3).
#include "k1_launch.cuh"

// launch geometry: two 192-thread CTAs per SM (12 warps at 166 registers).  4,194,304 systems: 123.5 ms against 126.9 ms for one
// 384-thread CTA (the CTA synchronises once per cycle; two half-size CTAs wait less) and 146.0 ms for one 512-thread CTA
// (128 registers, spills) -- GPU call 18 of round 2.
#ifndef K1_ROB_BLOCK
#define K1_ROB_BLOCK 192
#endif
#ifndef K1_ROB_MINB
#define K1_ROB_MINB 2
#endif

namespace brdae {

int k1_robertson(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    if (dense_mass) return k1_launch<Robertson, MassDense<3>, K1_ROB_BLOCK, K1_ROB_MINB, true>(a, sms, s, g, geom_only);
    return k1_launch<Robertson, Robertson::Mass, K1_ROB_BLOCK, K1_ROB_MINB, true>(a, sms, s, g, geom_only);
}

}  // namespace brdae
