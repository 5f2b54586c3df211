// k4_chain.cu -- kernel and launcher of K4, the lane-per-rope chain kernel (chain_k4.cuh).
#include <cuda_runtime.h>

#include <cstdlib>

#include "chain_k4.cuh"
#include "kernels.cuh"

namespace brdae {

namespace {

constexpr int K4_BLOCK = 32;             // one warp per CTA: the warps of an SM run different blocks of the step at the same time
constexpr size_t kHeaderBytes = 256;
size_t align8(size_t x) { return (x + 7) / 8 * 8; }

__global__ void __launch_bounds__(K4_BLOCK) k4_kernel(const __grid_constant__ SolveArgs a, double *state) {
    // [warp][node][slot][lane]: the state of the 32 ropes of a warp is one contiguous region
    const size_t warp = blockIdx.x;
    k4::Rope<32> R{state + warp * ((size_t)(a.n / 5) * k4::SPN * 32) + threadIdx.x};
    k4::k4_lane_loop<32>(a, R, true);
}

}  // namespace

// lane slots used for B ropes: one per rope up to a cap of about 48 GB of state
long long k4_slots(int n, int64_t B) {
    const long long per_slot = (long long)(n / 5) * k4::SPN * 8;
    long long cap = (48LL << 30) / per_slot;
    cap = cap / K4_BLOCK * K4_BLOCK;
    if (cap < K4_BLOCK) cap = K4_BLOCK;
    if (const char *e = std::getenv("BRDAE_K4_MAX_SLOTS")) {      // tests: fewer lane slots than ropes, so that lanes are refilled
        const long long v = std::atoll(e) / K4_BLOCK * K4_BLOCK;
        if (v >= K4_BLOCK && v < cap) cap = v;
    }
    long long G = (B + K4_BLOCK - 1) / K4_BLOCK * K4_BLOCK;
    if (G < K4_BLOCK) G = K4_BLOCK;
    return G < cap ? G : cap;
}

// workspace after the 256-byte work counter: atol[N] | var_index[N] | state [G/32][n/5][99][32]
size_t k4_scratch_bytes(int n, int64_t B) {
    const long long G = k4_slots(n, B);
    return align8(sizeof(double) * n) + align8(sizeof(int32_t) * n) + (size_t)(n / 5) * k4::SPN * G * sizeof(double);
}

int k4_chain(bool dense_mass, const SolveArgs &a_in, const brdae_batch *b, int sms, cudaStream_t stream, LaunchGeom *g, bool geom_only) {
    (void)sms;
    const int N = a_in.n;
    if (N < 10 || N % 5 != 0) return BRDAE_ERR_BAD_ARGUMENT;
    const long long G = k4_slots(N, a_in.B);
    if (g) { g->grid = (int)(G / K4_BLOCK); g->block = K4_BLOCK; g->smem = 0; g->scratch_bytes = 0; }
    if (geom_only) return BRDAE_OK;
    if (dense_mass || a_in.has_jac_const) return BRDAE_ERR_UNSUPPORTED;      // the block elimination is the chain's own M and J
    SolveArgs a = a_in;
    if (const char *e = std::getenv("BRDAE_NEWTON_REPS")) { const int v = std::atoi(e); if (v >= 1 && v <= 16) a.newton_reps = v; }
    char *ws = reinterpret_cast<char *>(a.work_counter) + kHeaderBytes;
    double *atol_d = reinterpret_cast<double *>(ws);
    int32_t *vidx_d = reinterpret_cast<int32_t *>(ws + align8(sizeof(double) * N));
    double *state = reinterpret_cast<double *>(ws + align8(sizeof(double) * N) + align8(sizeof(int32_t) * N));
    if (cudaMemcpyAsync(atol_d, b->atol, sizeof(double) * N, cudaMemcpyHostToDevice, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    if (b->var_index) {
        if (cudaMemcpyAsync(vidx_d, b->var_index, sizeof(int32_t) * N, cudaMemcpyHostToDevice, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    } else if (cudaMemsetAsync(vidx_d, 0, sizeof(int32_t) * N, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    a.atol_dev = atol_d; a.var_index_dev = vidx_d;
    k4_kernel<<<(int)(G / K4_BLOCK), K4_BLOCK, 0, stream>>>(a, state);
    return cudaPeekAtLastError() == cudaSuccess ? BRDAE_OK : BRDAE_ERR_CUDA;
}

}  // namespace brdae
