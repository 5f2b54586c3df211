// k1_extra.cu -- K1 instantiations for the reference's remaining known-answer problems:
// Robertson ODE formulation, Jay 1995 index-3 DAE, the polynomial index-2 DAE of the dense-output
// validation and the 7-variable linear ODE of the mass-matrix check.
#include "k1_launch.cuh"

namespace brdae {

template <class Prob, int BLOCK>
static int go(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    if (dense_mass) return k1_launch<Prob, MassDense<Prob::N>, BLOCK, 1, true>(a, sms, s, g, geom_only);
    return k1_launch<Prob, typename Prob::Mass, BLOCK, 1, true>(a, sms, s, g, geom_only);
}

int k1_extra(int problem, bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only) {
    switch (problem) {
    case BRDAE_ROBERTSON_ODE: return go<RobertsonODE, 384>(dense_mass, a, sms, s, g, geom_only);
    case BRDAE_JAY: return go<Jay, 256>(dense_mass, a, sms, s, g, geom_only);
    case BRDAE_POLYDAE2: return go<PolyDAE2, 256>(dense_mass, a, sms, s, g, geom_only);
    case BRDAE_LINEAR7: return go<Linear7, 128>(dense_mass, a, sms, s, g, geom_only);
    }
    return BRDAE_ERR_UNKNOWN_PROBLEM;
}

}  // namespace brdae
