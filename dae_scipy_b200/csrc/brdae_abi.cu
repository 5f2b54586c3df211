// brdae_abi.cu -- extern "C" entry points of libbrdae.so (see include/brdae.h).
// Host side only marshals: validates the batch, derives the uniform quantities the reference's
// constructor derives (radauDAE.py:263-422: newton_tol :322, NMAX_BAD :400-410, h_abs :307-315),
// packs them into SolveArgs and launches the kernel that matches (problem, mass kind).
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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
    case BRDAE_ERR_NONFINITE_Y0: return "all components of the initial state y0 must be finite";
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
    if (const ProblemInfo *u = find_problem(BRDAE_USER))   // plugin build: "user" or the functor's own name
        if (std::strcmp(name, "user") == 0 || std::strcmp(name, u->name) == 0) {
            if (problem) *problem = u->id;
            if (n) *n = u->n;
            if (n_params) *n_params = u->n_params;
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
    if (b->problem == BRDAE_NPENDULUM) {
        const bool fd = o && o->jac_finite_differences && !b->jac_const;      // finite differences: the band kernel
        bytes += (!fd && npendulum_uses_k4(b->n, b->n_systems)) ? k4_scratch_bytes(b->n, b->n_systems, sm_count())
                                                                  : k2_scratch_bytes(b->n, sm_count(), b->n_systems, fd);
    } else if (is_large_problem(b->problem)) {      // user band problem
        bytes += k2_user_scratch_bytes(b->n, sm_count(), b->n_systems, o && o->jac_finite_differences && !b->jac_const);
    } else {      // K1: per-CTA slice of cold per-lane state (kernels that keep it in shared memory need none)
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

// Row-per-system run of the K1 kernels: the batch arrays of `b` are [B][n] / [B][p] / [B][T][n] / [B][9] device
// buffers, and the kernel signals the completion of every `chunk_size` systems (see SolveArgs).
struct RowsRun { int32_t *chunk_done; volatile int32_t *chunk_flags; int64_t chunk_size; const int32_t *input_ready; const volatile int32_t *abort_flag; };

static int solve_impl(const brdae_batch *b, const brdae_options *o, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream, const RowsRun *rows) {
    int rc = validate(b, o);
    if (rc != BRDAE_OK) return rc;
    if (b->n_systems == 0) return BRDAE_OK;
    if (!workspace || workspace_bytes < brdae_workspace_bytes(b, o)) return BRDAE_ERR_WORKSPACE_TOO_SMALL;
    SolveArgs a;
    fill_common(a, b, o, workspace);
    if (rows) {
        if (is_large_problem(b->problem)) return BRDAE_ERR_UNSUPPORTED;       // the chain / band kernels are structure-of-arrays only
        const int64_t n = b->n, T = b->n_t_eval;
        a.y0_sc = 1; a.y0_sb = n; a.par_sc = 1; a.par_sb = b->n_params;
        a.ye_st = n; a.ye_sc = 1; a.ye_sb = n * T;      // the n unknowns of an output point are one 8n-byte run
        a.yf_sc = 1; a.yf_sb = n; a.cn_sq = 1; a.cn_sb = BRDAE_NCOUNTERS;
        a.chunk_done = rows->chunk_done; a.chunk_flags = rows->chunk_flags; a.chunk_size = rows->chunk_size;
        a.input_ready = rows->input_ready; a.abort_flag = rows->abort_flag;
    }
    if (cudaMemsetAsync(workspace, 0, kCounterBytes, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    rc = launch_solver(b->problem, b->mass != nullptr, a, b, sm_count(), stream);
    if (rc != BRDAE_OK) return rc;
    if (cudaGetLastError() != cudaSuccess) return BRDAE_ERR_CUDA;
    return BRDAE_OK;
}

int brdae_solve(const brdae_batch *b, const brdae_options *o, void *workspace, size_t workspace_bytes,
                void *cuda_stream) {
    return solve_impl(b, o, workspace, workspace_bytes, static_cast<cudaStream_t>(cuda_stream), nullptr);
}

#define CK(x) do { if ((x) != cudaSuccess) { rc = BRDAE_ERR_CUDA; goto done; } } while (0)

// Host-buffer form: stream-ordered allocation from the device's default memory pool (kept warm
// between calls by a high release threshold, so repeated calls do not pay cudaMalloc/cudaFree),
// H2D of the inputs, the solve, D2H of every result, one synchronisation at the end.  Pinned host
// buffers make the copies run at PCIe speed; pageable ones work but are staged by the driver.
int brdae_solve_host(const brdae_batch *hb, const brdae_options *o) {
    int rc = validate(hb, o);
    if (rc != BRDAE_OK) return rc;
    if (hb->dense_cap > 0 || hb->n_events > 0 || hb->report_cap > 0) return BRDAE_ERR_UNSUPPORTED;   // the per-step record, events and reports are device-buffer features
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
// y0 [B][n], params [B][p] in; y_eval [B][T][n], y_final [B][n], counters [B][9] out.  The
// rows <-> SoA permutations run on the device (layout.cu), so the host never transposes.
// Large ensembles are pipelined in chunks over two compute streams, one copy-out stream and three sets of
// device buffers (see the comment at the chunk loop).  A row-per-system slice is contiguous in host memory,
// so every copy stays one memcpy.
namespace {
struct RowsChunkBufs {
    double *stage = nullptr, *y0 = nullptr, *par = nullptr, *ye = nullptr, *ye_rows = nullptr, *tf = nullptr, *yf = nullptr,
           *yf_rows = nullptr;
    int32_t *ne = nullptr, *st = nullptr, *cn = nullptr, *cn_rows = nullptr;
    void *ws = nullptr;
    cudaEvent_t done = nullptr, freed = nullptr;
};
}  // namespace

// Thread-per-system problems: ONE launch over the whole ensemble, directly on row-per-system device buffers
// (the kernels take element strides), and a copy-out pipelined behind the kernel.  Systems are handed out in
// index order, so the ensemble completes roughly front to back: the kernel counts retired systems per chunk
// and raises a flag in host-mapped memory when a chunk is complete; this thread waits for the flags in order
// and queues each chunk's device-to-host copies on a second stream while the kernel keeps integrating the
// rest.  Chunking the LAUNCH instead (the chain-kernel path below) costs the refill tail of a persistent
// kernel per chunk -- about one trajectory's duration each, 21 ms over 8 chunks of the 1M-pendulum run.
static int host_rows_single_launch(const brdae_batch *hb, const brdae_options *o) {
    int rc = BRDAE_OK;
    const bool timing = std::getenv("BRDAE_ROWS_TIMING") != nullptr;      // tuning aid: phase times on stderr
    const auto h0 = std::chrono::steady_clock::now();
    auto hms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count(); };
    double t_setup = 0, t_launch = 0, t_inq = 0, t_flags[32] = {0}, t_kernel_done = 0;
    bool serial_in = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    const int64_t B = hb->n_systems;
    const int n = hb->n, T = hb->n_t_eval, np = hb->n_params;
    int nchunks = (int)(B / 65536);
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 32) nchunks = 32;
    const int64_t Bc = (B + nchunks - 1) / nchunks;
    nchunks = (int)((B + Bc - 1) / Bc);
    double *dy0 = nullptr, *dpar = nullptr, *dte = nullptr, *dye = nullptr, *dtf = nullptr, *dyf = nullptr;
    int32_t *dne = nullptr, *dst = nullptr, *dcn = nullptr, *dchunk = nullptr;
    volatile int32_t *flags = nullptr;
    void *ws = nullptr;
    cudaStream_t s = nullptr, s_copy = nullptr, s_in = nullptr;
    cudaEvent_t ready = nullptr;
    size_t wsb = 0;
    int dev = 0;
    cudaMemPool_t pool;
    CK(cudaGetDevice(&dev));
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    CK(cudaHostAlloc((void **)&flags, sizeof(int32_t) * 64, cudaHostAllocMapped));   // [0,32) completion flags, [32] the constant 1
    for (int k = 0; k < 64; ++k) flags[k] = (k == 32) ? 1 : 0;
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    if (timing) for (auto &e : ev) CK(cudaEventCreate(&e));
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    wsb = brdae_workspace_bytes(hb, o);
    CK(cudaMallocAsync(&dy0, sizeof(double) * n * B, s));
    if (np > 0) CK(cudaMallocAsync(&dpar, sizeof(double) * np * B, s));
    if (T > 0) {
        CK(cudaMallocAsync(&dte, sizeof(double) * T, s));
        CK(cudaMallocAsync(&dye, sizeof(double) * (size_t)T * n * B, s));
    }
    CK(cudaMallocAsync(&dtf, sizeof(double) * B, s));
    CK(cudaMallocAsync(&dyf, sizeof(double) * n * B, s));
    CK(cudaMallocAsync(&dne, sizeof(int32_t) * B, s));
    CK(cudaMallocAsync(&dst, sizeof(int32_t) * B, s));
    CK(cudaMallocAsync(&dcn, sizeof(int32_t) * BRDAE_NCOUNTERS * B, s));
    CK(cudaMallocAsync(&dchunk, sizeof(int32_t) * 64, s));           // [0,32) retired-system counts, [32,64) input-ready flags
    CK(cudaMallocAsync(&ws, wsb, s));
    CK(cudaMemsetAsync(dchunk, 0, sizeof(int32_t) * 64, s));
    if (T > 0) CK(cudaMemcpyAsync(dte, hb->t_eval, sizeof(double) * T, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ready, s));
    t_setup = hms();
    if (timing) CK(cudaEventRecord(ev[0], s));
    // The kernel is launched FIRST and its lanes wait for the chunk of the system they are handed; the copy-in
    // then streams chunk by chunk on its own stream, each chunk followed by its ready flag (from pageable host
    // memory cudaMemcpyAsync stages synchronously, so this thread is busy copying while the kernel already runs).
    // Under an injected CUDA tool (ncu, compute-sanitizer: CUDA_INJECTION64_PATH) a launch may not return before the
    // kernel has ended, and the copies it waits for would never be queued: copy in first, on the kernel's stream.
    // The same holds for CUDA_LAUNCH_BLOCKING=1 and for debuggers.  Should a launch block in a mode not recognised here, the
    // lanes' wait is bounded and the call returns BRDAE_ERR_CUDA instead of hanging (radau_k1.cuh, `input_lost`).
    {
        const char *lb = std::getenv("CUDA_LAUNCH_BLOCKING");
        serial_in = std::getenv("BRDAE_ROWS_SERIAL_IN") != nullptr       // A/B switch for tuning
                    || std::getenv("CUDA_INJECTION64_PATH") != nullptr || std::getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr
                    || std::getenv("CUDBG_USE_LEGACY_DEBUGGER") != nullptr || std::getenv("CUDA_DEBUGGER_SOFTWARE_PREEMPTION") != nullptr
                    || (lb && lb[0] != '\0' && lb[0] != '0');
    }
    if (serial_in) {
        CK(cudaMemcpyAsync(dy0, hb->y0, sizeof(double) * n * B, cudaMemcpyHostToDevice, s));
        if (np > 0) CK(cudaMemcpyAsync(dpar, hb->params, sizeof(double) * np * B, cudaMemcpyHostToDevice, s));
    }
    {
        brdae_batch d = *hb;
        d.y0 = dy0; d.params = dpar; d.t_eval = dte; d.y_eval = dye; d.t_final = dtf; d.y_final = dyf;
        d.n_eval = dne; d.status = dst; d.counters = dcn;
        RowsRun rr{dchunk, flags, Bc, serial_in ? nullptr : dchunk + 32, flags + 33};
        if (timing) CK(cudaEventRecord(ev[1], s));
        rc = solve_impl(&d, o, ws, wsb, s, &rr);
        if (rc != BRDAE_OK) goto done;
        if (timing) CK(cudaEventRecord(ev[2], s));
        t_launch = hms();
    }
    if (!serial_in) {
        CK(cudaStreamWaitEvent(s_in, ready, 0));
        for (int k = 0; k < nchunks; ++k) {
            const int64_t b0 = (int64_t)k * Bc, bn = (B - b0 < Bc) ? B - b0 : Bc;
            CK(cudaMemcpyAsync(dy0 + b0 * n, hb->y0 + b0 * n, sizeof(double) * n * bn, cudaMemcpyHostToDevice, s_in));
            if (np > 0) CK(cudaMemcpyAsync(dpar + b0 * np, hb->params + b0 * np, sizeof(double) * np * bn, cudaMemcpyHostToDevice, s_in));
            CK(cudaMemcpyAsync(dchunk + 32 + k, (const void *)(flags + 32), sizeof(int32_t), cudaMemcpyHostToDevice, s_in));
        }
    }
    t_inq = hms();
    for (int k = 0; k < nchunks; ++k) {
        const int64_t b0 = (int64_t)k * Bc, bn = (B - b0 < Bc) ? B - b0 : Bc;
        // wait for the chunk (or for the kernel to have ended, e.g. with an error)
        while (flags[k] == 0) {
            if (cudaStreamQuery(s) != cudaErrorNotReady) break;
        }
        if (flags[k] == 0) CK(cudaStreamSynchronize(s));          // kernel over: every result is in place
        t_flags[k] = hms();
        if (T > 0) CK(cudaMemcpyAsync(hb->y_eval + (size_t)b0 * n * T, dye + (size_t)b0 * n * T, sizeof(double) * (size_t)T * n * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->y_final + b0 * n, dyf + b0 * n, sizeof(double) * n * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->counters + b0 * BRDAE_NCOUNTERS, dcn + b0 * BRDAE_NCOUNTERS, sizeof(int32_t) * BRDAE_NCOUNTERS * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->t_final + b0, dtf + b0, sizeof(double) * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->n_eval + b0, dne + b0, sizeof(int32_t) * bn, cudaMemcpyDeviceToHost, s_copy));
        CK(cudaMemcpyAsync(hb->status + b0, dst + b0, sizeof(int32_t) * bn, cudaMemcpyDeviceToHost, s_copy));
    }
    if (cudaGetLastError() != cudaSuccess) rc = BRDAE_ERR_CUDA;
    {   // words 1 and 2 of the workspace header: did any lane give up waiting for its input?  does y0 hold NaN / inf?  (the
        // finiteness check scipy makes on the host, base.py:14-17, runs on the device copy behind the kernel: every chunk
        // has arrived by then)
        unsigned long long flags2[2] = {0, 0};
        count_nonfinite(dy0, (int64_t)n * B, reinterpret_cast<unsigned long long *>(static_cast<char *>(ws) + 16), s);
        CK(cudaMemcpyAsync(flags2, static_cast<char *>(ws) + 8, sizeof(flags2), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (!serial_in && flags2[0] != 0) rc = BRDAE_ERR_CUDA;
        else if (flags2[1] != 0) rc = BRDAE_ERR_NONFINITE_Y0;
    }
    if (timing) {
        CK(cudaEventRecord(ev[3], s_copy));
        CK(cudaStreamSynchronize(s)); t_kernel_done = hms();
        CK(cudaStreamSynchronize(s_copy));
        const double t_all = hms();
        float h2d = 0, kern = 0;
        cudaEventElapsedTime(&h2d, ev[0], ev[1]); cudaEventElapsedTime(&kern, ev[1], ev[2]);
        std::fprintf(stderr, "[brdae rows %s] setup %.2f ms, launch returned %.2f, copy-in queued %.2f, ev0-ev1 %.2f (device), kernel %.2f (device), kernel done at %.2f, "
                             "all copies done at %.2f; flags at", serial_in ? "serial-in" : "pipelined-in", t_setup, t_launch, t_inq, h2d, kern, t_kernel_done, t_all);
        for (int k = 0; k < nchunks; ++k) std::fprintf(stderr, " %.1f", t_flags[k]);
        std::fprintf(stderr, "\n");
    }
done:
    // an error between the launch and the last input copy leaves lanes waiting for chunks that will never arrive: tell them
    if (rc != BRDAE_OK && flags) flags[33] = 1;
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (s_in && cudaStreamSynchronize(s_in) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
    if (s_copy && cudaStreamSynchronize(s_copy) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
    if (s) {
        if (cudaStreamSynchronize(s) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
        void *bufs[] = {dy0, dpar, dte, dye, dtf, dyf, dne, dst, dcn, dchunk, ws};
        for (void *q : bufs) if (q) cudaFreeAsync(q, s);
        if (cudaStreamSynchronize(s) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
        cudaStreamDestroy(s);
    }
    if (s_copy) cudaStreamDestroy(s_copy);
    if (s_in) cudaStreamDestroy(s_in);
    if (ready) cudaEventDestroy(ready);
    if (flags) cudaFreeHost((void *)flags);
    return rc;
}

int brdae_solve_host_rows(const brdae_batch *hb, const brdae_options *o) {
    int rc = validate(hb, o);
    if (rc != BRDAE_OK) return rc;
    if (hb->dense_cap > 0 || hb->n_events > 0 || hb->report_cap > 0) return BRDAE_ERR_UNSUPPORTED;
    const int64_t B = hb->n_systems;
    if (B == 0) return BRDAE_OK;
    if (!is_large_problem(hb->problem) && !std::getenv("BRDAE_ROWS_CHUNKED")) return host_rows_single_launch(hb, o);
    const int n = hb->n, T = hb->n_t_eval, np = hb->n_params;
    const int64_t in_rows = (int64_t)(n > np ? n : np);
    // Chunked pipeline: about 128k systems per chunk (at most 16 chunks), three buffer sets, two compute
    // streams and one copy-out stream.  Chunk c+1's H2D, permutation and kernel are queued on the other
    // compute stream while chunk c still runs, so its CTAs move onto the SMs as the persistent CTAs of chunk c
    // run out of work (the refill tail of one launch is filled by the next), and the D2H of chunk c overlaps
    // the solve of chunks c+1, c+2.  Exposed transfer time is one chunk's H2D and one chunk's D2H.
    constexpr int MAXSETS = 3;
    int64_t chunk_systems = 131072;
    if (const char *e = std::getenv("BRDAE_ROWS_CHUNK")) { const long long v = std::atoll(e); if (v >= 1024) chunk_systems = v; }   // tuning
    int nchunks = (int)(B / chunk_systems);
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 16) nchunks = 16;
    const int64_t Bc = (B + nchunks - 1) / nchunks;
    const int nsets = nchunks < MAXSETS ? nchunks : MAXSETS;
    RowsChunkBufs cb[MAXSETS];
    double *dte = nullptr;
    cudaStream_t sc[2] = {nullptr, nullptr}, s_copy = nullptr;
    cudaEvent_t ready = nullptr;
    size_t wsb = 0;
    int dev = 0;
    cudaMemPool_t pool;
    CK(cudaGetDevice(&dev));
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    CK(cudaStreamCreateWithFlags(&sc[0], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sc[1], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    {
        cudaStream_t s = sc[0];           // every allocation is ordered on the first compute stream
        if (T > 0) {
            CK(cudaMallocAsync(&dte, sizeof(double) * T, s));
            CK(cudaMemcpyAsync(dte, hb->t_eval, sizeof(double) * T, cudaMemcpyHostToDevice, s));
        }
        brdae_batch probe = *hb;
        probe.n_systems = Bc;
        wsb = brdae_workspace_bytes(&probe, o);
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
        CK(cudaEventRecord(ready, s));
        CK(cudaStreamWaitEvent(sc[1], ready, 0));
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        const int64_t b0 = (int64_t)ch * Bc;
        const int64_t bn = (B - b0 < Bc) ? B - b0 : Bc;
        if (bn <= 0) break;
        RowsChunkBufs &c = cb[ch % nsets];
        cudaStream_t s = sc[ch & 1];
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
        if (T > 0) permute_f64(c.ye, c.ye_rows, (int64_t)T * n, bn, bn, (int64_t)T * n, 1, 0, 0, s);   // [T*n][bn] -> [bn][T][n]
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
    for (cudaStream_t q : {sc[1], s_copy})
        if (q && cudaStreamSynchronize(q) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
    if (sc[0]) {
        cudaStream_t s = sc[0];
        if (cudaStreamSynchronize(s) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
        for (int k = 0; k < MAXSETS; ++k) {
            RowsChunkBufs &c = cb[k];
            void *bufs[] = {c.stage, c.y0, c.par, c.ye, c.ye_rows, c.tf, c.yf, c.yf_rows, c.ne, c.st, c.cn, c.cn_rows, c.ws};
            for (void *q : bufs) if (q) cudaFreeAsync(q, s);
        }
        if (dte) cudaFreeAsync(dte, s);
        if (cudaStreamSynchronize(s) != cudaSuccess && rc == BRDAE_OK) rc = BRDAE_ERR_CUDA;
        for (int k = 0; k < MAXSETS; ++k) {
            if (cb[k].done) cudaEventDestroy(cb[k].done);
            if (cb[k].freed) cudaEventDestroy(cb[k].freed);
        }
    }
    if (ready) cudaEventDestroy(ready);
    for (cudaStream_t q : {sc[0], sc[1], s_copy}) if (q) cudaStreamDestroy(q);
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
