// k1_robertson.cu -- K1 instantiations for the Robertson DAE (n = 3).
#include "k1_launch.cuh"

#ifndef K1_ROB_BLOCK
#define K1_ROB_BLOCK 384
#endif
#ifndef K1_ROB_MINB
#define K1_ROB_MINB 1
#endif

namespace brdae {

int k1_robertson(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    if (dense_mass) return k1_launch<Robertson, MassDense<3>, K1_ROB_BLOCK, K1_ROB_MINB, true>(a, sms, s, g, geom_only);
    return k1_launch<Robertson, Robertson::Mass, K1_ROB_BLOCK, K1_ROB_MINB, true>(a, sms, s, g, geom_only);
}

}  // namespace brdae
