// brdae_abi.cu -- extern "C" entry points of libbrdae.so (see include/brdae.h).
// Host side only marshals: validates the batch, derives the uniform quantities the reference's
// constructor derives (radauDAE.py:263-422: newton_tol :322, NMAX_BAD :400-410, h_abs :307-315),
// packs them into SolveArgs and launches the kernel that matches (problem, mass kind).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>

#include "host_args.hpp"
#include "kernels.cuh"

using namespace brdae;

namespace {

int sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms;
}

}  // namespace

extern "C" {

int brdae_abi_version(void) { return BRDAE_ABI_VERSION; }

const char *brdae_strerror(int code) {
    switch (code) {
    case BRDAE_OK: return "ok";
    case BRDAE_ERR_BAD_ARGUMENT: return "bad argument";
    case BRDAE_ERR_UNKNOWN_PROBLEM: return "unknown problem id";
    case BRDAE_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case BRDAE_ERR_CUDA: return "CUDA runtime error";
    case BRDAE_ERR_UNSUPPORTED: return "unsupported configuration";
    }
    return "unknown error code";
}

int brdae_problem_lookup(const char *name, int32_t *problem, int32_t *n, int32_t *n_params) {
    if (!name) return BRDAE_ERR_BAD_ARGUMENT;
    for (const auto &p : kProblems)
        if (std::strcmp(p.name, name) == 0) {
            if (problem) *problem = p.id;
            if (n) *n = p.n;
            if (n_params) *n_params = p.n_params;
            return BRDAE_OK;
        }
    return BRDAE_ERR_UNKNOWN_PROBLEM;
}

void brdae_default_options(brdae_options *o) {
    if (!o) return;
    o->first_step = -1.0; o->max_step = INFINITY; o->newton_tol = -1.0;
    o->factor_on_non_convergence = 0.5; o->safety_factor = 0.9; o->min_factor = 0.2; o->max_factor = 10.0;
    o->jac_recompute_factor = 1e-3; o->max_newton_ite = 6; o->max_bad_ite = -1;
    o->scale_residuals = 1; o->scale_newton_norm = 1; o->scale_error = 1; o->zero_algebraic_error = 0;
    o->always_2nd_estimate = 1; o->predictive_controller = 1; o->predictive_newton_stop = 1;
    o->extrapolated_guess = 1; o->all_newton_iterations = 0; o->constant_dt = 0; o->nmax_step = 0;
}

size_t brdae_workspace_bytes(const brdae_batch *b, const brdae_options *) {
    if (!b) return 0;
    size_t bytes = kCounterBytes;
    if (b->problem == BRDAE_NPENDULUM) bytes += k2_scratch_bytes(b->n, sm_count());
    return bytes;
}

int brdae_launch_info(const brdae_batch *b, int32_t *grid, int32_t *block, int32_t *smem_bytes) {
    if (!b) return BRDAE_ERR_BAD_ARGUMENT;
    LaunchGeom g;
    int rc = launch_geometry(b->problem, b->mass != nullptr, b->n, b->n_systems, sm_count(), &g);
    if (rc != BRDAE_OK) return rc;
    if (grid) *grid = g.grid;
    if (block) *block = g.block;
    if (smem_bytes) *smem_bytes = g.smem;
    return BRDAE_OK;
}

int brdae_solve(const brdae_batch *b, const brdae_options *o, void *workspace, size_t workspace_bytes,
                void *cuda_stream) {
    int rc = validate(b, o);
    if (rc != BRDAE_OK) return rc;
    if (b->n_systems == 0) return BRDAE_OK;
    if (!workspace || workspace_bytes < brdae_workspace_bytes(b, o)) return BRDAE_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    SolveArgs a;
    fill_common(a, b, o, workspace);
    if (cudaMemsetAsync(workspace, 0, kCounterBytes, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    rc = launch_solver(b->problem, b->mass != nullptr, a, b, sm_count(), stream);
    if (rc != BRDAE_OK) return rc;
    if (cudaGetLastError() != cudaSuccess) return BRDAE_ERR_CUDA;
    return BRDAE_OK;
}

#define CK(x) do { if ((x) != cudaSuccess) { rc = BRDAE_ERR_CUDA; goto done; } } while (0)

int brdae_solve_host(const brdae_batch *hb, const brdae_options *o) {
    int rc = validate(hb, o);
    if (rc != BRDAE_OK) return rc;
    const int64_t B = hb->n_systems;
    if (B == 0) return BRDAE_OK;
    const int n = hb->n, T = hb->n_t_eval, np = hb->n_params;
    brdae_batch d = *hb;
    double *dy0 = nullptr, *dpar = nullptr, *dte = nullptr, *dye = nullptr, *dtf = nullptr, *dyf = nullptr;
    int32_t *dne = nullptr, *dst = nullptr, *dcn = nullptr;
    void *ws = nullptr;
    cudaStream_t s = nullptr;
    size_t wsb = 0;
    CK(cudaStreamCreate(&s));
    CK(cudaMalloc(&dy0, sizeof(double) * n * B));
    CK(cudaMemcpyAsync(dy0, hb->y0, sizeof(double) * n * B, cudaMemcpyHostToDevice, s));
    if (np > 0) {
        CK(cudaMalloc(&dpar, sizeof(double) * np * B));
        CK(cudaMemcpyAsync(dpar, hb->params, sizeof(double) * np * B, cudaMemcpyHostToDevice, s));
    }
    if (T > 0) {
        CK(cudaMalloc(&dte, sizeof(double) * T));
        CK(cudaMemcpyAsync(dte, hb->t_eval, sizeof(double) * T, cudaMemcpyHostToDevice, s));
        CK(cudaMalloc(&dye, sizeof(double) * (size_t)T * n * B));
    }
    CK(cudaMalloc(&dtf, sizeof(double) * B));
    CK(cudaMalloc(&dyf, sizeof(double) * n * B));
    CK(cudaMalloc(&dne, sizeof(int32_t) * B));
    CK(cudaMalloc(&dst, sizeof(int32_t) * B));
    CK(cudaMalloc(&dcn, sizeof(int32_t) * BRDAE_NCOUNTERS * B));
    d.y0 = dy0; d.params = dpar; d.t_eval = dte; d.y_eval = dye; d.t_final = dtf; d.y_final = dyf;
    d.n_eval = dne; d.status = dst; d.counters = dcn;
    wsb = brdae_workspace_bytes(&d, o);
    CK(cudaMalloc(&ws, wsb));
    rc = brdae_solve(&d, o, ws, wsb, s);
    if (rc != BRDAE_OK) goto done;
    if (T > 0) CK(cudaMemcpyAsync(hb->y_eval, dye, sizeof(double) * (size_t)T * n * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->t_final, dtf, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->y_final, dyf, sizeof(double) * n * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->n_eval, dne, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->status, dst, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->counters, dcn, sizeof(int32_t) * BRDAE_NCOUNTERS * B, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
done:
    cudaFree(dy0); cudaFree(dpar); cudaFree(dte); cudaFree(dye); cudaFree(dtf); cudaFree(dyf);
    cudaFree(dne); cudaFree(dst); cudaFree(dcn); cudaFree(ws);
    if (s) cudaStreamDestroy(s);
    return rc;
}

int brdae_fp64_peak_probe(double ms_target, double *flops, void *cuda_stream) {
    if (!flops) return BRDAE_ERR_BAD_ARGUMENT;
    return fp64_peak_probe(ms_target, flops, static_cast<cudaStream_t>(cuda_stream), sm_count());
}

}  // extern "C"
