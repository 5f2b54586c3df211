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
size_t align256(size_t x) { return (x + 255) / 256 * 256; }       // the state starts on a line boundary (the header is 256 bytes)

// RPW ropes per warp (lanes RPW..31 idle), at least MINB warps resident per SM.  The kernel is bound by the latency of
// its global-memory accesses (one warp per SM measured 63 % long-scoreboard stalls and 3 % of the issue slots used,
// profiles/r02_k4_ncu_summary.md): the ropes are the only source of parallelism, so with fewer ropes than the machine has
// lanes they are spread over MORE WARPS rather than more lanes -- the idle lanes cost nothing while the schedulers have
// no other warp to issue from, and a warp with fewer ropes loses fewer lanes to divergence between its ropes.
template <int RPW, int MINB>
__global__ void __launch_bounds__(K4_BLOCK, MINB) k4_kernel(const __grid_constant__ SolveArgs a, double *state) {
    // [warp][node][slot][rope of the warp]: the state of the ropes of a warp is one contiguous region
    extern __shared__ __align__(128) unsigned char k4_smem[];      // the warp's ring of staged node blocks (chain_k4.cuh: Stager)
    const size_t warp = blockIdx.x;
    const bool have = threadIdx.x < RPW;
    k4::Rope<RPW> R{state + warp * ((size_t)(a.n / 5) * k4::SPN * RPW) + (have ? threadIdx.x : 0)};
    constexpr bool STAGED = BRDAE_K4_STAGED && RPW >= 2;            // bulk copies move multiples of 16 bytes
    k4::Stager<RPW, STAGED> S;
    if constexpr (STAGED) S.init(k4_smem);
    k4::k4_lane_loop<RPW>(a, R, have, S);
}
template <int RPW>
constexpr int k4_smem_bytes() { return (BRDAE_K4_STAGED && RPW >= 2) ? k4::STAGE_SLOTS * RPW * 8 * 3 + 64 : 0; }

struct K4Geom { int rpw, minb; long long warps; };
int k4_smem_for(int rpw);

int env_int(const char *name, int lo, int hi, int dflt) {
    if (const char *e = std::getenv(name)) { const int v = std::atoi(e); if (v >= lo && v <= hi) return v; }
    return dflt;
}

// ropes per warp and warps for a batch of B ropes on `sms` SMs: as many warps as stay resident (MINB per SM), ropes per
// warp the smallest power of two that fits the batch into them; lanes are refilled from the work counter beyond that
K4Geom k4_geometry(int n, int64_t B, int sms) {
    const int nsm = sms > 0 ? sms : 148;
    K4Geom g;
    const int mb = env_int("BRDAE_K4_MINB", 1, 32, 8);      // 252 registers, no spills: measured best (profiles/r02_k4_ncu_summary.md, "Geometry")
    g.minb = mb >= 24 ? 32 : (mb >= 12 ? 16 : 8);
    // warps that stay resident on an SM: the register budget (minb) and the staged node ring in shared memory (k4_smem_for)
    auto resident_for = [&](int rpw) {
        const int by_smem = (int)((227 * 1024) / (k4_smem_for(rpw) + 1024));
        return (long long)nsm * (g.minb < by_smem ? g.minb : by_smem);
    };
    int rpw = 1;
    while (rpw < 32 && (long long)rpw * resident_for(rpw) < B) rpw *= 2;
    // With the node stream staged in shared memory a warp no longer waits on global loads, and a warp instruction costs the
    // same for 32 ropes as for 8: as many ropes per warp as leave about two warps per SM (16,384 ropes, n = 20: 32 ropes per
    // warp 171 ms, 16: 258 ms, 8: 418 ms; the unstaged kernel preferred 16: 317 ms -- profiles/r02_k4_ncu_summary.md).
    if (BRDAE_K4_STAGED)
        while (rpw < 32 && B / (2 * rpw) >= 2LL * nsm) rpw *= 2;
    rpw = env_int("BRDAE_K4_RPW", 1, 32, rpw);
    int p2 = 1;
    while (p2 < rpw) p2 *= 2;
    g.rpw = p2;
    const long long resident = resident_for(g.rpw);
    long long warps = (B + g.rpw - 1) / g.rpw;
    if (warps > resident) warps = resident;
    // memory cap (about 48 GB of state) and the test hook that forces refills
    const long long per_slot = (long long)(n / 5) * k4::SPN * 8;
    long long cap = (48LL << 30) / per_slot;
    cap = (long long)env_int("BRDAE_K4_MAX_SLOTS", 1, 1 << 30, (int)(cap > (1 << 30) ? (1 << 30) : cap));
    if (warps * g.rpw > cap) warps = cap / g.rpw;
    if (warps < 1) warps = 1;
    g.warps = warps;
    return g;
}

template <int RPW, int MINB>
void k4_launch_one(const K4Geom &g, const SolveArgs &a, double *state, cudaStream_t stream) {
    constexpr int smem = k4_smem_bytes<RPW>();
    if (smem > 48 * 1024) cudaFuncSetAttribute(k4_kernel<RPW, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k4_kernel<RPW, MINB><<<(int)g.warps, K4_BLOCK, smem, stream>>>(a, state);
}
template <int MINB>
cudaError_t k4_launch(const K4Geom &g, const SolveArgs &a, double *state, cudaStream_t stream) {
    switch (g.rpw) {
    case 1: k4_launch_one<1, MINB>(g, a, state, stream); break;
    case 2: k4_launch_one<2, MINB>(g, a, state, stream); break;
    case 4: k4_launch_one<4, MINB>(g, a, state, stream); break;
    case 8: k4_launch_one<8, MINB>(g, a, state, stream); break;
    case 16: k4_launch_one<16, MINB>(g, a, state, stream); break;
    default: k4_launch_one<32, MINB>(g, a, state, stream); break;
    }
    return cudaPeekAtLastError();
}

int k4_smem_for(int rpw) {
    switch (rpw) {
    case 1: return k4_smem_bytes<1>();
    case 2: return k4_smem_bytes<2>();
    case 4: return k4_smem_bytes<4>();
    case 8: return k4_smem_bytes<8>();
    case 16: return k4_smem_bytes<16>();
    default: return k4_smem_bytes<32>();
    }
}

}  // namespace

// workspace after the 256-byte work counter: atol[N] | var_index[N] | state [warps][n/5][99][ropes per warp]
size_t k4_scratch_bytes(int n, int64_t B, int sms) {
    const K4Geom g = k4_geometry(n, B, sms);
    return align256(align8(sizeof(double) * n) + align8(sizeof(int32_t) * n)) + (size_t)(n / 5) * k4::SPN * g.warps * g.rpw * sizeof(double);
}

int k4_chain(bool dense_mass, const SolveArgs &a_in, const brdae_batch *b, int sms, cudaStream_t stream, LaunchGeom *g, bool geom_only) {
    const int N = a_in.n;
    if (N < 10 || N % 5 != 0) return BRDAE_ERR_BAD_ARGUMENT;
    const K4Geom kg = k4_geometry(N, a_in.B, sms);
    if (g) { g->grid = (int)kg.warps; g->block = K4_BLOCK; g->smem = k4_smem_for(kg.rpw); g->scratch_bytes = 0; }
    if (geom_only) return BRDAE_OK;
    if (dense_mass || a_in.has_jac_const) return BRDAE_ERR_UNSUPPORTED;      // the block elimination is the chain's own M and J
    SolveArgs a = a_in;
    if (const char *e = std::getenv("BRDAE_NEWTON_REPS")) { const int v = std::atoi(e); if (v >= 1 && v <= 16) a.newton_reps = v; }
    a.newton_vote = env_int("BRDAE_K4_VOTE", 0, 32, 1);     // keep iterating until at most 1/32 of the live ropes still do
    char *ws = reinterpret_cast<char *>(a.work_counter) + kHeaderBytes;
    double *atol_d = reinterpret_cast<double *>(ws);
    int32_t *vidx_d = reinterpret_cast<int32_t *>(ws + align8(sizeof(double) * N));
    double *state = reinterpret_cast<double *>(ws + align256(align8(sizeof(double) * N) + align8(sizeof(int32_t) * N)));
    if (cudaMemcpyAsync(atol_d, b->atol, sizeof(double) * N, cudaMemcpyHostToDevice, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    if (b->var_index) {
        if (cudaMemcpyAsync(vidx_d, b->var_index, sizeof(int32_t) * N, cudaMemcpyHostToDevice, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    } else if (cudaMemsetAsync(vidx_d, 0, sizeof(int32_t) * N, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    a.atol_dev = atol_d; a.var_index_dev = vidx_d;
    const cudaError_t err = (kg.minb >= 32) ? k4_launch<32>(kg, a, state, stream)
                            : (kg.minb >= 16) ? k4_launch<16>(kg, a, state, stream) : k4_launch<8>(kg, a, state, stream);
    return err == cudaSuccess ? BRDAE_OK : BRDAE_ERR_CUDA;
}

}  // namespace brdae
