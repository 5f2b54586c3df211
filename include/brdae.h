/* brdae.h -- C ABI of the B200-native batched RadauDAE solver (libbrdae.so).
 *
 * What this boundary replaces.  The reference (laurent90git/DAE-Scipy) has no native FFI: its
 * plug-in interface is scipy's OdeSolver subclass protocol, selected with
 *     solve_ivp(fun, t_span, y0, method=RadauDAE, t_eval=..., **options)
 * (reference README.md:19-21; class RadauDAE at scipyDAE/radauDAE.py:126; constructor keywords
 * :263-278; step loop :515-791; driver with t_eval slicing :1079-1157 and scipy ivp.py:659-732).
 * One call of `brdae_solve` stands for B such solve_ivp calls -- one per system of the ensemble --
 * with the problem's RHS and analytic Jacobian compiled in as device functors (or, with
 * jac_finite_differences, the reference's default finite-difference Jacobian of that RHS)
 * (reference closures: tests/pendulum.py:130-148, tests/robertson.py:80-87,
 * tests/transistor_amplifier.py:48-56, tests/recursive_pendulum.py:93-118).
 *
 * Conventions: plain pointers and sizes only; the caller owns every buffer; the library keeps
 * no global mutable state; every entry returns an int error code (0 = ok) and never throws;
 * `brdae_solve` is asynchronous on the given CUDA stream.  Pointers marked [dev] are device
 * pointers, [host] are host pointers that are read before the call returns.
 * Batch arrays are structure-of-arrays: element (row, sys) lives at row*B + sys.
 */
#ifndef BRDAE_H
#define BRDAE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRDAE_ABI_VERSION 4
#define BRDAE_NMAX_SMALL 8 /* largest state size of the thread-per-system kernels */
#define BRDAE_MAX_EVENTS 4 /* event functions located on the device per call */

/* problems (device functors); ids are stable */
enum {
    BRDAE_PENDULUM3 = 0, /* tests/pendulum.py:129-148  params (m, g, r0)            n=5 */
    BRDAE_PENDULUM2 = 1, /* tests/pendulum.py:150-166                              n=5 */
    BRDAE_PENDULUM1 = 2, /* tests/pendulum.py:168-184                              n=5 */
    BRDAE_ROBERTSON = 3, /* tests/robertson.py:78-87   params (k1, k2, k3)          n=3 */
    BRDAE_TRANSISTOR = 4, /* tests/transistor_amplifier.py:33-56
                             params (C1,C2,C3,R0,R1,R2,R3,R4,R5,Ub,Ua,w)           n=5 */
    BRDAE_NPENDULUM = 5, /* tests/recursive_pendulum.py:93-127, no params, n = 5*nodes */
    BRDAE_ROBERTSON_ODE = 6, /* tests/robertson.py:69-77 (ODE form), params (k1,k2,k3)   n=3 */
    BRDAE_JAY = 7,       /* tests/jay_dae.py:22-68, params (nonlinear flag)             n=5 */
    BRDAE_POLYDAE2 = 8,  /* tests/validation_dense_output.py:23-59 (p=2), params a4..a0 n=2 */
    BRDAE_LINEAR7 = 9,   /* tests/linear_ODE_with_mass_matrix.py:32-39, params c[7]     n=7 */
    /* The one problem of a PLUGIN build of this library: the `fun` / `jac` closures a user of the reference
     * hands to solve_ivp (radauDAE.py:263-266, 464-513), written as a device functor (n <= BRDAE_NMAX_SMALL)
     * and compiled into a copy of the library (csrc/k1_user.cu; dae_scipy_b200/plugin.py shows the recipe).
     * Its name, n and n_params are reported by brdae_problem_lookup(); the stock library answers
     * BRDAE_ERR_UNKNOWN_PROBLEM for it. */
    BRDAE_USER = 100
};

/* per-system status, same meaning as OdeResult.status (radauDAE.py:986-989, 1087-1091) */
enum {
    BRDAE_STATUS_FINISHED = 0,
    BRDAE_STATUS_EVENT = 1,           /* "A termination event occurred." (scipy ivp.py:15-16; radauDAE.py:1110-1129) */
    BRDAE_STATUS_TOO_SMALL_STEP = -1, /* OdeSolver.TOO_SMALL_STEP, radauDAE.py:565-568 */
    BRDAE_STATUS_LU_FAILED = -2,      /* radauDAE.py:611-613 (never produced: see DESIGN.md) */
    BRDAE_STATUS_NMAX_STEP = 2        /* solve_ivp_custom nmax_step, radauDAE.py:1081-1084 */
};

/* rows of the counters array [BRDAE_NCOUNTERS][B] (radauDAE.py:347-352 + scipy base nfev/njev/nlu) */
enum {
    BRDAE_C_NFEV = 0, BRDAE_C_NJEV, BRDAE_C_NLU, BRDAE_C_NSTEP, BRDAE_C_NACCPT, BRDAE_C_NREJCT,
    BRDAE_C_NFAILED, BRDAE_C_NLUSOLVE, BRDAE_C_NLUSOLVE_ERR, BRDAE_NCOUNTERS
};

/* error codes */
enum {
    BRDAE_OK = 0, BRDAE_ERR_BAD_ARGUMENT = 1, BRDAE_ERR_UNKNOWN_PROBLEM = 2,
    BRDAE_ERR_WORKSPACE_TOO_SMALL = 3, BRDAE_ERR_CUDA = 4, BRDAE_ERR_UNSUPPORTED = 5,
    BRDAE_ERR_NONFINITE_Y0 = 6      /* brdae_solve_host_rows: y0 holds NaN or inf (scipy raises ValueError for it, base.py:14-17) */
};

/* POD mirror of the RadauDAE constructor keywords (radauDAE.py:263-278) plus solve_ivp_custom's
 * nmax_step (:992).  Fill with brdae_default_options() first. */
typedef struct brdae_options {
    double first_step;  /* <=0: reference default with a mass matrix, min(1e-6,|t_bound-t0|) (:313) */
    double max_step;    /* +inf */
    double newton_tol;  /* <=0: max(10 eps/rtol, min(0.03, sqrt(rtol))) (:322) */
    double factor_on_non_convergence; /* 0.5 */
    double safety_factor;             /* 0.9 */
    double min_factor;                /* 0.2 */
    double max_factor;                /* 10 */
    double jac_recompute_factor;      /* jacobianRecomputeFactor = 1e-3 */
    int32_t max_newton_ite;           /* 6 */
    int32_t max_bad_ite;              /* <0: 1 if any var_index>1 else 0 (:406-410) */
    int32_t scale_residuals;          /* 1 */
    int32_t scale_newton_norm;        /* 1 */
    int32_t scale_error;              /* 1 */
    int32_t zero_algebraic_error;     /* 0 */
    int32_t always_2nd_estimate;      /* bAlwaysApply2ndEstimate = 1 */
    int32_t predictive_controller;    /* bUsePredictiveController = 1 */
    int32_t predictive_newton_stop;   /* bUsePredictiveNewtonStoppingCriterion = 1 */
    int32_t extrapolated_guess;       /* bUseExtrapolatedGuess = 1 */
    int32_t all_newton_iterations;    /* bPerformAllNewtonIterations = 0 */
    int32_t constant_dt;              /* 0 */
    int32_t jac_finite_differences;   /* 0; 1 = jac=None of the reference: scipy num_jac forward differences
                                         with adaptive column factors (radauDAE.py:468-481); ignored when
                                         jac_const is given; thread-per-system problems only */
    int32_t reserved;                 /* 0 */
    int64_t nmax_step;                /* <=0: unlimited */
} brdae_options;

/* one ensemble = B independent systems of the same problem and option set */
typedef struct brdae_batch {
    int32_t problem;          /* BRDAE_* problem id */
    int32_t n;                /* state size (5*nodes for BRDAE_NPENDULUM) */
    int32_t n_params;         /* rows of `params` */
    int32_t n_t_eval;         /* T, may be 0 */
    int64_t n_systems;        /* B */
    const double *y0;         /* [dev]  [n][B] initial states */
    const double *params;     /* [dev]  [n_params][B], may be NULL when n_params == 0 */
    const double *t_eval;     /* [dev]  [T] strictly monotone in the direction of integration */
    const double *mass;       /* [host] [n][n] row-major shared mass matrix, NULL = the problem's own */
    const double *jac_const;  /* [host] [n][n] constant Jacobian (radauDAE.py:502-511), NULL = analytic functor */
    const int32_t *var_index; /* [host] [n] algebraic index per unknown (0..3), NULL = all 0 (:389-398) */
    const double *atol;       /* [host] [n] */
    double rtol;
    double t0, t_bound;
    double *y_eval;           /* [dev]  [T][n][B]; NaN where a failed system did not get to */
    int32_t *n_eval;          /* [dev]  [B] number of t_eval points written */
    double *t_final;          /* [dev]  [B] */
    double *y_final;          /* [dev]  [n][B] */
    int32_t *status;          /* [dev]  [B] BRDAE_STATUS_* */
    int32_t *counters;        /* [dev]  [BRDAE_NCOUNTERS][B] */
    /* Optional record of every accepted step: what solve_ivp(dense_output=True) keeps as OdeSolution
     * (radauDAE.py:951-982, 1097-1134) and, with t_eval=None, as OdeResult.t / .y.  Step k of a system
     * ends at dense_t[k] with state dense_y[k]; inside it y(t) = y_old + Q (x, x^2, x^3)^T with
     * x = (t - t_old)/(dense_t[k] - t_old), (t_old, y_old) the end of step k-1 (or t0, y0).  A system
     * records its first min(naccpt, dense_cap) steps.  All NULL / 0 = off.  brdae_solve only. */
    int32_t dense_cap;        /* K */
    int32_t reserved0;
    double *dense_t;          /* [dev]  [K][B] */
    double *dense_y;          /* [dev]  [K][n][B] */
    double *dense_q;          /* [dev]  [K][n][3][B]: Q[c][j] of step k at ((k*n + c)*3 + j)*B + sys */
    /* Optional EVENTS located on the device, with the semantics of solve_ivp(events=...) (scipy ivp.py
     * prepare_events / find_active_events / handle_events, which radauDAE.py:1059-1129 wraps), for event
     * functions that are linear in the state:  g_e(t, y) = sum_j event_coef[e][j] * y_j - event_level[e].
     * After every accepted step an event whose g changed sign in the allowed direction (event_direction[e]:
     * < 0 only + to -, > 0 only - to +, 0 both) is located on the step's cubic interpolant; its time and state are
     * recorded (the first event_cap occurrences per event and system; event_count counts all of them) and, if
     * event_terminal[e] = k > 0, the k-th occurrence ends that system there: status BRDAE_STATUS_EVENT,
     * t_final / y_final = the event point, t_eval points beyond it are not produced.  n_events = 0 = off.
     * Thread-per-system problems, brdae_solve only. */
    int32_t n_events;         /* E <= BRDAE_MAX_EVENTS */
    int32_t event_cap;        /* C */
    const double *event_coef;        /* [host] [E][n] */
    const double *event_level;       /* [host] [E] */
    const int32_t *event_direction;  /* [host] [E] */
    const int32_t *event_terminal;   /* [host] [E] */
    int32_t *event_count;     /* [dev]  [E][B] */
    double *event_t;          /* [dev]  [E][C][B]; entries at or beyond the count are not written */
    double *event_y;          /* [dev]  [E][C][n][B], may be NULL */
    const int32_t *event_functor;    /* [host] [E] or NULL: k >= 0 = event e is the k-th event function g_k(t, y) of the problem
                                        functor itself (plugin builds: UserProblem::event, any nonlinear function; coef and
                                        level of that event are ignored), -1 = the linear form above */
    /* Optional PER-ATTEMPT REPORT, the reference's `bReport=True` (`RadauDAE._report`, radauDAE.py:424-443, codes :24-28):
     * one record per factorisation (-10), Jacobian update after a non-converged Newton with a stale Jacobian (-11),
     * non-converged attempt (5), rejected step (2) and accepted step (0), in the order the reference appends them.  A
     * system records its first report_cap events; report_count counts all of them.  report_cap = 0 = off.  brdae_solve only. */
    int32_t report_cap;       /* R */
    int32_t reserved1;
    int32_t *report_count;    /* [dev]  [B] */
    int32_t *report_code;     /* [dev]  [R][B] */
    int32_t *report_it;       /* [dev]  [R][2][B]: Newton iterations and bad iterations of the attempt (-1 for codes -10, -11) */
    double *report_val;       /* [dev]  [R][4][B]: t, h, err1, err2 (the first / second error estimate: -1 where the reference
                                        passes none, NaN for an estimate that was not computed) */
} brdae_batch;

int brdae_abi_version(void);
const char *brdae_strerror(int code);

/* name -> id, state size (0 when it depends on the call, i.e. npendulum) and parameter count */
int brdae_problem_lookup(const char *name, int32_t *problem, int32_t *n, int32_t *n_params);

void brdae_default_options(brdae_options *opts);

/* bytes of device scratch `brdae_solve` needs for this batch (work queue + per-system scratch) */
size_t brdae_workspace_bytes(const brdae_batch *batch, const brdae_options *opts);

/* Integrate the ensemble on the current device, asynchronously on `cuda_stream`
 * (a cudaStream_t passed as void*; NULL = default stream). */
int brdae_solve(const brdae_batch *batch, const brdae_options *opts, void *workspace,
                size_t workspace_bytes, void *cuda_stream);

/* Same call with HOST buffers for every [dev] pointer of `batch`: allocates device memory,
 * copies in, solves, copies out and synchronises.  This is the form a solve_ivp-style caller with
 * numpy arrays binds to. */
int brdae_solve_host(const brdae_batch *host_batch, const brdae_options *opts);

/* Host-buffer call in the row-per-system layout a per-system solve_ivp user already holds:
 * y0 [B][n], params [B][n_params] in; y_eval [B][T][n] (the transpose of OdeResult.y per system: the n
 * unknowns of one output point are contiguous, which is how a lane writes them), y_final [B][n],
 * counters [B][BRDAE_NCOUNTERS] out; n_eval, t_final, status [B].  The permutations to and from
 * the kernels' structure-of-arrays layout are done on the device. */
int brdae_solve_host_rows(const brdae_batch *host_batch_rows, const brdae_options *opts);

/* Page-locked host buffers for callers of brdae_solve_host (cudaMallocHost / cudaFreeHost). */
int brdae_alloc_pinned(size_t bytes, void **ptr);
int brdae_free_pinned(void *ptr);

/* Launch geometry chosen for a batch, for reporting (grid, block, dynamic shared memory bytes). */
int brdae_launch_info(const brdae_batch *batch, int32_t *grid, int32_t *block, int32_t *smem_bytes);

/* FP64 FMA throughput probe: runs a dependent-chain DFMA kernel for about `ms_target`
 * milliseconds on the current device and returns the achieved FLOP/s (2 per FMA) in *flops.
 * Used as the measured FP64 roofline denominator. */
int brdae_fp64_peak_probe(double ms_target, double *flops, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* BRDAE_H */
