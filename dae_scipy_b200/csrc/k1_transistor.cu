// k1_transistor.cu -- K1 instantiations for the transistor amplifier (n = 5, singular
// non-diagonal mass matrix built from each system's capacitances).
#include "k1_launch.cuh"

namespace brdae {

int k1_transistor(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    if (dense_mass) return k1_launch<Transistor, MassDense<5>, 256, 1, true>(a, sms, s, g, geom_only);
    return k1_launch<Transistor, Transistor::Mass, 256, 1, true>(a, sms, s, g, geom_only);
}

}  // namespace brdae
