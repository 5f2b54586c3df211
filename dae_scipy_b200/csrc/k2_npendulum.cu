// k2_npendulum.cu -- placeholder translation unit until the CTA-per-system banded kernel lands.
#include "kernels.cuh"

namespace brdae {

size_t k2_scratch_bytes(int, int) { return 0; }

int k2_npendulum(bool, const SolveArgs &, const brdae_batch *, int, cudaStream_t, LaunchGeom *g, bool) {
    if (g) { g->grid = 0; g->block = 0; g->smem = 0; }
    return BRDAE_ERR_UNSUPPORTED;
}

}  // namespace brdae
