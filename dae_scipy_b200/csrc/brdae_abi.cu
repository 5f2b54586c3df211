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
    o->jac_finite_differences = 0; o->reserved = 0;
}

size_t brdae_workspace_bytes(const brdae_batch *b, const brdae_options *o) {
    if (!b) return 0;
    size_t bytes = kCounterBytes;
    if (b->problem == BRDAE_NPENDULUM)
        bytes += npendulum_uses_k3(b->n, b->n_systems) ? k3_scratch_bytes(b->n, b->n_systems) : k2_scratch_bytes(b->n, sm_count(), b->n_systems);
    else {      // K1: per-CTA slice of cold per-lane state (kernels that keep it in shared memory need none)
        LaunchGeom g{};
        const bool fd = o && o->jac_finite_differences && !b->jac_const;
        if (launch_geometry(b->problem, b->mass != nullptr, fd, b->n, b->n_systems, sm_count(), &g) == BRDAE_OK) bytes += g.scratch_bytes;
    }
    return bytes;
}

int brdae_launch_info(const brdae_batch *b, int32_t *grid, int32_t *block, int32_t *smem_bytes) {
    if (!b) return BRDAE_ERR_BAD_ARGUMENT;
    LaunchGeom g;
    int rc = launch_geometry(b->problem, b->mass != nullptr, false, b->n, b->n_systems, sm_count(), &g);
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

// Host-buffer form: stream-ordered allocation from the device's default memory pool (kept warm
// between calls by a high release threshold, so repeated calls do not pay cudaMalloc/cudaFree),
// H2D of the inputs, the solve, D2H of every result, one synchronisation at the end.  Pinned host
// buffers make the copies run at PCIe speed; pageable ones work but are staged by the driver.
int brdae_solve_host(const brdae_batch *hb, const brdae_options *o) {
    int rc = validate(hb, o);
    if (rc != BRDAE_OK) return rc;
    if (hb->dense_cap > 0) return BRDAE_ERR_UNSUPPORTED;   // the per-step record is a device-buffer feature
    const int64_t B = hb->n_systems;
    if (B == 0) return BRDAE_OK;
    const int n = hb->n, T = hb->n_t_eval, np = hb->n_params;
    brdae_batch d = *hb;
    double *dy0 = nullptr, *dpar = nullptr, *dte = nullptr, *dye = nullptr, *dtf = nullptr, *dyf = nullptr;
    int32_t *dne = nullptr, *dst = nullptr, *dcn = nullptr;
    void *ws = nullptr;
    cudaStream_t s = nullptr;
    size_t wsb = 0;
    int dev = 0;
    cudaMemPool_t pool;
    CK(cudaGetDevice(&dev));
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaMallocAsync(&dy0, sizeof(double) * n * B, s));
    CK(cudaMemcpyAsync(dy0, hb->y0, sizeof(double) * n * B, cudaMemcpyHostToDevice, s));
    if (np > 0) {
        CK(cudaMallocAsync(&dpar, sizeof(double) * np * B, s));
        CK(cudaMemcpyAsync(dpar, hb->params, sizeof(double) * np * B, cudaMemcpyHostToDevice, s));
    }
    if (T > 0) {
        CK(cudaMallocAsync(&dte, sizeof(double) * T, s));
        CK(cudaMemcpyAsync(dte, hb->t_eval, sizeof(double) * T, cudaMemcpyHostToDevice, s));
        CK(cudaMallocAsync(&dye, sizeof(double) * (size_t)T * n * B, s));
    }
    CK(cudaMallocAsync(&dtf, sizeof(double) * B, s));
    CK(cudaMallocAsync(&dyf, sizeof(double) * n * B, s));
    CK(cudaMallocAsync(&dne, sizeof(int32_t) * B, s));
    CK(cudaMallocAsync(&dst, sizeof(int32_t) * B, s));
    CK(cudaMallocAsync(&dcn, sizeof(int32_t) * BRDAE_NCOUNTERS * B, s));
    d.y0 = dy0; d.params = dpar; d.t_eval = dte; d.y_eval = dye; d.t_final = dtf; d.y_final = dyf;
    d.n_eval = dne; d.status = dst; d.counters = dcn;
    wsb = brdae_workspace_bytes(&d, o);
    CK(cudaMallocAsync(&ws, wsb, s));
    rc = brdae_solve(&d, o, ws, wsb, s);
    if (rc != BRDAE_OK) goto done;
    if (T > 0) CK(cudaMemcpyAsync(hb->y_eval, dye, sizeof(double) * (size_t)T * n * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->t_final, dtf, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->y_final, dyf, sizeof(double) * n * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->n_eval, dne, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->status, dst, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hb->counters, dcn, sizeof(int32_t) * BRDAE_NCOUNTERS * B, cudaMemcpyDeviceToHost, s));
done:
    if (s) {
        void *bufs[] = {dy0, dpar, dte, dye, dtf, dyf, dne, dst, dcn, ws};
        for (void *q : bufs) if (q) cudaFreeAsync(q, s);
        if (cudaStreamSynchronize(s) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
        cudaStreamDestroy(s);
    }
    return rc;
}

// Row-per-system form (the layout a caller coming from per-system solve_ivp calls holds):
// y0 [B][n], params [B][p] in; y_eval [B][n][T], y_final [B][n], counters [B][9] out.  The
// rows <-> SoA permutations run on the device (layout.cu), so the host never transposes.
// Large ensembles are pipelined in chunks over two streams and two sets of device buffers: the
// device-to-host copy of chunk c (the bulk of the PCIe traffic) overlaps the integration of chunk
// c+1.  A row-per-system slice is contiguous in host memory, so every copy stays one memcpy.
namespace {
struct RowsChunkBufs {
    double *stage = nullptr, *y0 = nullptr, *par = nullptr, *ye = nullptr, *ye_rows = nullptr, *tf = nullptr, *yf = nullptr,
           *yf_rows = nullptr;
    int32_t *ne = nullptr, *st = nullptr, *cn = nullptr, *cn_rows = nullptr;
    void *ws = nullptr;
    cudaEvent_t done = nullptr, freed = nullptr;
};
}  // namespace

int brdae_solve_host_rows(const brdae_batch *hb, const brdae_options *o) {
    int rc = validate(hb, o);
    if (rc != BRDAE_OK) return rc;
    if (hb->dense_cap > 0) return BRDAE_ERR_UNSUPPORTED;
    const int64_t B = hb->n_systems;
    if (B == 0) return BRDAE_OK;
    const int n = hb->n, T = hb->n_t_eval, np = hb->n_params;
    const int64_t in_rows = (int64_t)(n > np ? n : np);
    // chunking: about a quarter of a million systems per chunk, at most 8 chunks
    int nchunks = (int)(B / 262144);
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 8) nchunks = 8;
    const int64_t Bc = (B + nchunks - 1) / nchunks;
    const int nsets = nchunks > 1 ? 2 : 1;
    RowsChunkBufs cb[2];
    double *dte = nullptr;
    cudaStream_t s = nullptr, s_copy = nullptr;
    size_t wsb = 0;
    int dev = 0;
    cudaMemPool_t pool;
    CK(cudaGetDevice(&dev));
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    if (T > 0) {
        CK(cudaMallocAsync(&dte, sizeof(double) * T, s));
        CK(cudaMemcpyAsync(dte, hb->t_eval, sizeof(double) * T, cudaMemcpyHostToDevice, s));
    }
    {
        brdae_batch probe = *hb;
        probe.n_systems = Bc;
        wsb = brdae_workspace_bytes(&probe, o);
    }
    for (int k = 0; k < nsets; ++k) {
        RowsChunkBufs &c = cb[k];
        CK(cudaMallocAsync(&c.stage, sizeof(double) * in_rows * Bc, s));
        CK(cudaMallocAsync(&c.y0, sizeof(double) * n * Bc, s));
        if (np > 0) CK(cudaMallocAsync(&c.par, sizeof(double) * np * Bc, s));
        if (T > 0) {
            CK(cudaMallocAsync(&c.ye, sizeof(double) * (size_t)T * n * Bc, s));
            CK(cudaMallocAsync(&c.ye_rows, sizeof(double) * (size_t)T * n * Bc, s));
        }
        CK(cudaMallocAsync(&c.tf, sizeof(double) * Bc, s));
        CK(cudaMallocAsync(&c.yf, sizeof(double) * n * Bc, s));
        CK(cudaMallocAsync(&c.yf_rows, sizeof(double) * n * Bc, s));
        CK(cudaMallocAsync(&c.ne, sizeof(int32_t) * Bc, s));
        CK(cudaMallocAsync(&c.st, sizeof(int32_t) * Bc, s));
        CK(cudaMallocAsync(&c.cn, sizeof(int32_t) * BRDAE_NCOUNTERS * Bc, s));
        CK(cudaMallocAsync(&c.cn_rows, sizeof(int32_t) * BRDAE_NCOUNTERS * Bc, s));
        CK(cudaMallocAsync(&c.ws, wsb, s));
        CK(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.freed, cudaEventDisableTiming));
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        const int64_t b0 = (int64_t)ch * Bc;
        const int64_t bn = (B - b0 < Bc) ? B - b0 : Bc;
        if (bn <= 0) break;
        RowsChunkBufs &c = cb[ch % nsets];
        if (ch >= nsets) CK(cudaStreamWaitEvent(s, c.freed, 0));           // this set's results have left the device
        CK(cudaMemcpyAsync(c.stage, hb->y0 + b0 * n, sizeof(double) * n * bn, cudaMemcpyHostToDevice, s));
        permute_f64(c.stage, c.y0, bn, n, n, bn, 1, 0, 0, s);               // [bn][n] -> [n][bn]
        if (np > 0) {
            CK(cudaMemcpyAsync(c.stage, hb->params + b0 * np, sizeof(double) * np * bn, cudaMemcpyHostToDevice, s));
            permute_f64(c.stage, c.par, bn, np, np, bn, 1, 0, 0, s);
        }
        brdae_batch d = *hb;
        d.n_systems = bn;
        d.y0 = c.y0; d.params = c.par; d.t_eval = dte; d.y_eval = c.ye; d.t_final = c.tf; d.y_final = c.yf;
        d.n_eval = c.ne; d.status = c.st; d.counters = c.cn;
        rc = brdae_solve(&d, o, c.ws, wsb, s);
        if (rc != BRDAE_OK) goto done;
        if (T > 0) permute_f64(c.ye, c.ye_rows, T, bn, (int64_t)n * bn, (int64_t)n * T, n, bn, T, s);   // -> [bn][n][T]
        permute_f64(c.yf, c.yf_rows, n, bn, bn, n, 1, 0, 0, s);
        permute_i32(c.cn, c.cn_rows, BRDAE_NCOUNTERS, bn, bn, BRDAE_NCOUNTERS, 1, 0, 0, s);
        CK(cudaEventRecord(c.done, s));
        CK(cudaStreamWaitEvent(s_copy, c.done, 0));
        if (T > 0) CK(cudaMemcpyAsync(hb->y_eval + (size_t)b0 * n * T, c.ye_rows, sizeof(double) * (size_t)T * n * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->y_final + b0 * n, c.yf_rows, sizeof(double) * n * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->counters + b0 * BRDAE_NCOUNTERS, c.cn_rows, sizeof(int32_t) * BRDAE_NCOUNTERS * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->t_final + b0, c.tf, sizeof(double) * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->n_eval + b0, c.ne, sizeof(int32_t) * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->status + b0, c.st, sizeof(int32_t) * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaEventRecord(c.freed, s_copy));
    }
    if (cudaGetLastError() != cudaSuccess) rc = BRDAE_ERR_CUDA;
done:
    if (s_copy && cudaStreamSynchronize(s_copy) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
    if (s) {
        for (int k = 0; k < 2; ++k) {
            RowsChunkBufs &c = cb[k];
            void *bufs[] = {c.stage, c.y0, c.par, c.ye, c.ye_rows, c.tf, c.yf, c.yf_rows, c.ne, c.st, c.cn, c.cn_rows, c.ws};
            for (void *q : bufs) if (q) cudaFreeAsync(q, s);
        }
        if (dte) cudaFreeAsync(dte, s);
        if (cudaStreamSynchronize(s) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
        for (int k = 0; k < 2; ++k) {
            if (cb[k].done) cudaEventDestroy(cb[k].done);
            if (cb[k].freed) cudaEventDestroy(cb[k].freed);
        }
        cudaStreamDestroy(s);
    }
    if (s_copy) cudaStreamDestroy(s_copy);
    return rc;
}

// Page-locked host memory for callers without their own allocator (numpy users): copies from and
// to such buffers run at full PCIe rate and do not pay first-touch page faults on every call.
int brdae_alloc_pinned(size_t bytes, void **ptr) {
    if (!ptr) return BRDAE_ERR_BAD_ARGUMENT;
    return cudaMallocHost(ptr, bytes ? bytes : 1) == cudaSuccess ? BRDAE_OK : BRDAE_ERR_CUDA;
}
int brdae_free_pinned(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? BRDAE_OK : BRDAE_ERR_CUDA; }

int brdae_fp64_peak_probe(double ms_target, double *flops, void *cuda_stream) {
    if (!flops) return BRDAE_ERR_BAD_ARGUMENT;
    return fp64_peak_probe(ms_target, flops, static_cast<cudaStream_t>(cuda_stream), sm_count());
}

}  // extern "C"
