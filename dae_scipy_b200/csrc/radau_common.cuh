// radau_common.cuh -- shared declarations of the batched RadauDAE kernels (sm_100a).
//
// Radau IIA tableau transforms: radauDAE.py:11-48 (values as exact hex literals).
// Arithmetic contract: DESIGN.md section 4 -- the translation unit is compiled with
// -fmad=false, so the fma() calls written in the kernels are the only fused operations and
// every kernel can be compared bit for bit with the scalar restatement used by the tests.
#pragma once
#include <cmath>
#include <cstdint>

#include "../../include/brdae.h"

#if defined(__CUDACC__)
#define BRDAE_HD __device__ __forceinline__
#define BRDAE_D __device__ __forceinline__
#define BRDAE_CX __host__ __device__ constexpr
#else
#define BRDAE_HD inline
#define BRDAE_D inline
#define BRDAE_CX constexpr
#endif

namespace brdae {

constexpr double c_RC0 = 0x1.3d8b64657caeap-3, c_RC1 = 0x1.4a36c0803a6dfp-1;
constexpr double c_RE0 = -0x1.418fd8baffe05p+3, c_RE1 = 0x1.61d41b2d54580p+0, c_RE2 = -0x1.5555555555555p-2;
constexpr double c_MU_REAL = 0x1.d1a48d83e731dp+1;
constexpr double c_MU_CR = 0x1.572db93e0c672p+1, c_MU_CI = -0x1.86747f2c3fcb6p+1;
constexpr double c_T00 = 0x1.82d23845d677bp-4, c_T01 = -0x1.214a74c403bf7p-3, c_T02 = 0x1.ebff91a6d688ep-6;
constexpr double c_T10 = 0x1.0037de70aa1edp-2, c_T11 = 0x1.a20e91e20b551p-3, c_T12 = -0x1.8821fa2a36dedp-2;
constexpr double c_TI00 = 0x1.0b70201a79c46p+2, c_TI01 = 0x1.4f8c15da84e86p-2, c_TI02 = 0x1.0bf7ff59d56efp-1;
constexpr double c_TI10 = -0x1.0b70201a79c46p+2, c_TI11 = -0x1.4f8c15da84e86p-2, c_TI12 = 0x1.e810014c55221p-2;
constexpr double c_TI20 = 0x1.017885a24a7f2p-1, c_TI21 = -0x1.4934e6fcaa598p+1, c_TI22 = 0x1.312c0cf7bdfd1p-1;
constexpr double c_P00 = 0x1.418fd8baffe05p+3, c_P01 = -0x1.9a12ce7b30915p+4, c_P02 = 0x1.f295c43b61425p+3;
constexpr double c_P10 = -0x1.61d41b2d54580p+0, c_P11 = 0x1.497af24bb677ep+3, c_P12 = -0x1.1d406ee60becfp+3;
constexpr double c_P20 = 0x1.5555555555555p-2, c_P21 = -0x1.5555555555555p+1, c_P22 = 0x1.aaaaaaaaaaaabp+1;


// On the device the tableau constants are read from constant memory (LDCU.128 into uniform registers, two constants per
// instruction) instead of being written into the code as immediates, which ptxas materialised with two UMOVs per constant and
// use inside the rolled Newton loops (5.5 percent of the warp instructions of the pendulum kernel).  Same values, same bits.
// Measured (GPU call 14 of round 2, constant memory against immediates): pendulum 58.03 / 57.96 ms, Robertson 127.0 / 128.6 ms,
// transistor 134.9 / 135.5 ms, chain n = 20 161.1 / 165.1 ms -- the kernels are latency-bound, not issue-bound, so the saved
// instructions buy 0-2 percent.  BRDAE_CONST_BANK=0 restores the immediates (A/B builds).
#ifndef BRDAE_CONST_BANK
#define BRDAE_CONST_BANK 1
#endif
#if defined(__CUDACC__) && BRDAE_CONST_BANK
__constant__ double kTableau[32] = {c_RC0, c_RC1, c_RE0, c_RE1, c_RE2, c_MU_REAL, c_MU_CR, c_MU_CI, c_T00, c_T01, c_T02, c_T10, c_T11, c_T12, c_TI00, c_TI01, c_TI02, c_TI10, c_TI11, c_TI12, c_TI20, c_TI21, c_TI22, c_P00, c_P01, c_P02, c_P10, c_P11, c_P12, c_P20, c_P21, c_P22};
#if defined(__CUDA_ARCH__)
#define BRDAE_K(name, i) (::brdae::kTableau[i])
#else
#define BRDAE_K(name, i) (::brdae::name)
#endif
#else
#define BRDAE_K(name, i) (::brdae::name)
#endif
#define RC0 BRDAE_K(c_RC0, 0)
#define RC1 BRDAE_K(c_RC1, 1)
#define RE0 BRDAE_K(c_RE0, 2)
#define RE1 BRDAE_K(c_RE1, 3)
#define RE2 BRDAE_K(c_RE2, 4)
#define MU_REAL BRDAE_K(c_MU_REAL, 5)
#define MU_CR BRDAE_K(c_MU_CR, 6)
#define MU_CI BRDAE_K(c_MU_CI, 7)
#define T00 BRDAE_K(c_T00, 8)
#define T01 BRDAE_K(c_T01, 9)
#define T02 BRDAE_K(c_T02, 10)
#define T10 BRDAE_K(c_T10, 11)
#define T11 BRDAE_K(c_T11, 12)
#define T12 BRDAE_K(c_T12, 13)
#define TI00 BRDAE_K(c_TI00, 14)
#define TI01 BRDAE_K(c_TI01, 15)
#define TI02 BRDAE_K(c_TI02, 16)
#define TI10 BRDAE_K(c_TI10, 17)
#define TI11 BRDAE_K(c_TI11, 18)
#define TI12 BRDAE_K(c_TI12, 19)
#define TI20 BRDAE_K(c_TI20, 20)
#define TI21 BRDAE_K(c_TI21, 21)
#define TI22 BRDAE_K(c_TI22, 22)
#define P00 BRDAE_K(c_P00, 23)
#define P01 BRDAE_K(c_P01, 24)
#define P02 BRDAE_K(c_P02, 25)
#define P10 BRDAE_K(c_P10, 26)
#define P11 BRDAE_K(c_P11, 27)
#define P12 BRDAE_K(c_P12, 28)
#define P20 BRDAE_K(c_P20, 29)
#define P21 BRDAE_K(c_P21, 30)
#define P22 BRDAE_K(c_P22, 31)

// the same transforms as tables, for loops that are deliberately not unrolled
#if defined(__CUDACC__)
#define BRDAE_TABLE __constant__ const double
#else
#define BRDAE_TABLE static const double
#endif
BRDAE_TABLE kStageT[3][3] = {{c_T00, c_T01, c_T02}, {c_T10, c_T11, c_T12}, {1.0, 1.0, 0.0}};
BRDAE_TABLE kStageTI[3][3] = {{c_TI00, c_TI01, c_TI02}, {c_TI10, c_TI11, c_TI12}, {c_TI20, c_TI21, c_TI22}};
BRDAE_TABLE kStageC[3] = {c_RC0, c_RC1, 1.0};

// (T W)_i for one component: fma chain over the stages 0,1,2; T[2] = (1, 1, 0)
template <int I>
BRDAE_HD double stage_z(double w0, double w1, double w2) {
    if (I == 0) return fma(T02, w2, fma(T01, w1, T00 * w0));
    if (I == 1) return fma(T12, w2, fma(T11, w1, T10 * w0));
    return w0 + w1;
}
BRDAE_HD double stage_c(int i) { return i == 0 ? RC0 : (i == 1 ? RC1 : 1.0); }

// finite test that reads the same on host and device (false for NaN and +-inf)
BRDAE_HD bool is_finite(double x) { return fabs(x) <= 1.7976931348623157e308; }

BRDAE_HD double hpow(double h, int e) { return e == 0 ? 1.0 : (e == 1 ? h : h * h); }

// predict_factor, radauDAE.py:52-98; x**0.25 = sqrt(sqrt(x))
BRDAE_HD double gustafsson(double h_abs, double h_abs_old, double err, double err_old, bool hist,
                           bool predictive) {
    double mult = 1.0;
    if (predictive && hist && err != 0.0) mult = h_abs / h_abs_old * sqrt(sqrt(err_old / err));
    double base = 1.0 / sqrt(sqrt(err));
    return (mult < 1.0 ? mult : 1.0) * base;
}

// kernel-side view of one call: everything uniform across the ensemble lives here (constant bank)
struct SolveArgs {
    int64_t B;
    int32_t T;
    int32_t n;  // runtime state size (npendulum); equals Prob::N for the small problems
    const double *y0, *params, *t_eval;
    double *y_eval;
    int32_t *n_eval;
    double *t_final, *y_final;
    int32_t *status, *counters;
    // element strides of the batch arrays (K1): entry (row c, system s) of y0 lives at c*y0_sc + s*y0_sb, and so
    // on; y_eval entry (point e, row c, system s) at e*ye_st + c*ye_sc + s*ye_sb.  Structure-of-arrays
    // (brdae_solve) is sc = B, sb = 1; the row-per-system entry runs the kernel directly on [B][n] / [B][n][T].
    int64_t y0_sc, y0_sb, par_sc, par_sb, ye_st, ye_sc, ye_sb, yf_sc, yf_sb, cn_sq, cn_sb;
    // optional completion signalling for pipelined copy-out (row-per-system entry): systems are handed out in
    // index order, chunk k = systems [k*chunk_size, (k+1)*chunk_size); the lane that retires the last system
    // of a chunk raises chunk_flags[k] (host-mapped memory) after a system-wide fence
    int32_t *chunk_done;
    volatile int32_t *chunk_flags;
    int64_t chunk_size;
    // optional pipelined copy-in: input_ready[k] != 0 once y0/params of chunk k have arrived (written by a
    // host-to-device copy queued behind the chunk's data); a lane waits for its system's chunk before loading it
    const volatile int32_t *input_ready;
    // set by the host (mapped memory) when the call is abandoned while lanes may still wait for input_ready: waiting
    // lanes give up instead of spinning forever; a lane whose wait exceeds its bound raises work_counter[1] instead
    const volatile int32_t *abort_flag;
    double *dense_t, *dense_y, *dense_q;   // optional per-step record (brdae.h), dense_cap steps per system
    int32_t dense_cap;
    // optional device-side events (brdae.h): g_e(y) = ev_coef[e] . y - ev_level[e]
    int32_t n_events, ev_cap;
    int32_t ev_dir[BRDAE_MAX_EVENTS], ev_term[BRDAE_MAX_EVENTS], ev_functor[BRDAE_MAX_EVENTS];   // ev_functor: -1 = linear form
    double ev_level[BRDAE_MAX_EVENTS], ev_coef[BRDAE_MAX_EVENTS][BRDAE_NMAX_SMALL];
    int32_t *ev_count;
    double *ev_t, *ev_y;
    // optional per-attempt report (brdae.h: report_*), rep_cap records per system
    int32_t rep_cap;
    int32_t *rep_count, *rep_code, *rep_it;
    double *rep_val;
    unsigned long long *work_counter;
    double *scratch;  // per-CTA global scratch (K2)
    double t0, tb, dir, rtol;
    double h0, max_step, newton_tol, nonconv_factor, safety_factor, min_factor, max_factor, jac_recompute;
    double rsq3n, rsqn, rsq_nondiff;
    int64_t nmax_step;
    int32_t maxit, nmax_bad;
    int32_t scale_residuals, scale_newton_norm, scale_error, zero_algebraic_error;
    int32_t always_2nd, predictive_controller, predictive_newton, extrapolate, all_newton, constant_dt;
    int32_t has_jac_const;
    int32_t jac_fd;        // K1: finite-difference Jacobian (jac=None of the reference)
    int32_t newton_reps;   // K1: Newton passes per cycle (scheduling only, results do not depend on it)
    int32_t newton_vote;   // K4: after newton_reps passes, keep iterating while more than newton_vote/32 of the live ropes do (0: never)
    // small-problem (n <= BRDAE_NMAX_SMALL) uniform vectors; K2 reads its own from global memory
    double atol[BRDAE_NMAX_SMALL];
    int32_t var_index[BRDAE_NMAX_SMALL];
    double mass[BRDAE_NMAX_SMALL * BRDAE_NMAX_SMALL];
    double jac_const[BRDAE_NMAX_SMALL * BRDAE_NMAX_SMALL];
    // large-problem uniform vectors (device pointers, K2)
    const double *atol_dev;
    const int32_t *var_index_dev;
};

enum : int { CN_NFEV = 0, CN_NJEV, CN_NLU, CN_NSTEP, CN_NACC, CN_NREJ, CN_NFAIL, CN_NSOLVE, CN_NSOLVE_ERR, CN_COUNT };

// codes of the per-attempt report (radauDAE.py:24-28)
enum : int { RC_FACTORISATION = -10, RC_JACOBIAN_UPDATE = -11, RC_NON_CONV = 5, RC_ACCEPTED = 0, RC_REJECTED = 2 };
// append one record to system idx's report (the reference's RadauDAE._report, radauDAE.py:424-443)
BRDAE_HD void report_event(const SolveArgs &a, long long idx, int code, double t, double h, int newt, int nbad, double e1, double e2) {
    const int k = a.rep_count[idx];
    a.rep_count[idx] = k + 1;
    if (k >= a.rep_cap) return;
    a.rep_code[(int64_t)k * a.B + idx] = code;
    a.rep_it[((int64_t)k * 2 + 0) * a.B + idx] = newt; a.rep_it[((int64_t)k * 2 + 1) * a.B + idx] = nbad;
    a.rep_val[((int64_t)k * 4 + 0) * a.B + idx] = t; a.rep_val[((int64_t)k * 4 + 1) * a.B + idx] = h;
    a.rep_val[((int64_t)k * 4 + 2) * a.B + idx] = e1; a.rep_val[((int64_t)k * 4 + 3) * a.B + idx] = e2;
}

}  // namespace brdae
