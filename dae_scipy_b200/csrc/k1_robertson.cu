// k1_robertson.cu -- K1 instantiations for the Robertson DAE (n = 3).
#include "k1_launch.cuh"

namespace brdae {

int k1_robertson(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    if (dense_mass) return k1_launch<Robertson, MassDense<3>, 128, 4>(a, sms, s, g, geom_only);
    return k1_launch<Robertson, Robertson::Mass, 128, 4>(a, sms, s, g, geom_only);
}

}  // namespace brdae
