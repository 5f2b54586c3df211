/* ORACLE (test infrastructure, NOT product code) -- scalar C restatement of RadauDAE.
 *
 * One system at a time, plain C99, no vector tricks.  It restates the algorithm of the
 * reference `scipyDAE/radauDAE.py` (RadauDAE._step_impl :515-791, solve_collocation_system
 * :793-949, predict_factor :52-98, dense output :951-982, constructor :263-422) and of the
 * solve_ivp driver loop that slices t_eval (radauDAE.py:1079-1157 / scipy ivp.py:659-732),
 * for the dense branch with an analytic (or constant) Jacobian.
 *
 * Arithmetic contract ("the spec").  The reference's numbers come out
 * of numpy + OpenBLAS, whose rounding cannot be reproduced elsewhere (the
 * reference is not even bit-reproducible against itself across OpenBLAS kernel families).
 * This restatement therefore fixes ONE explicit evaluation order -- right-looking LU with
 * partial pivoting and reciprocal pivots, fma() where written, x**0.25 as sqrt(sqrt(x)),
 * divisions by the Newton/error scales and by h done as multiplications by a reciprocal, norms as
 * sqrt(sum)*rsqrt(size) -- so that an independent implementation of the same spec (the CUDA
 * kernels) can be compared with it BIT FOR BIT, while this file itself is pinned to the
 * reference statistically (tests/test_oracle_c.py: every t_eval value within the parity
 * tolerance of the reference-generated fixtures and identical accepted/rejected counts on
 * the well-conditioned configurations).  Compile with -ffp-contract=off: fma() calls are the
 * only fused operations.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the shared object built from this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- Radau IIA constants, radauDAE.py:11-48 (hex values: SURVEY 3.4) ---- */
static const double RC[3] = {0x1.3d8b64657caeap-3, 0x1.4a36c0803a6dfp-1, 1.0};
static const double RE[3] = {-0x1.418fd8baffe05p+3, 0x1.61d41b2d54580p+0, -0x1.5555555555555p-2};
static const double MU_REAL = 0x1.d1a48d83e731dp+1;
static const double MU_CR = 0x1.572db93e0c672p+1, MU_CI = -0x1.86747f2c3fcb6p+1;
static const double RT[3][3] = {
    {0x1.82d23845d677bp-4, -0x1.214a74c403bf7p-3, 0x1.ebff91a6d688ep-6},
    {0x1.0037de70aa1edp-2, 0x1.a20e91e20b551p-3, -0x1.8821fa2a36dedp-2},
    {1.0, 1.0, 0.0}};
static const double RTI[3][3] = {
    {0x1.0b70201a79c46p+2, 0x1.4f8c15da84e86p-2, 0x1.0bf7ff59d56efp-1},
    {-0x1.0b70201a79c46p+2, -0x1.4f8c15da84e86p-2, 0x1.e810014c55221p-2},
    {0x1.017885a24a7f2p-1, -0x1.4934e6fcaa598p+1, 0x1.312c0cf7bdfd1p-1}};
static const double RP[3][3] = {
    {0x1.418fd8baffe05p+3, -0x1.9a12ce7b30915p+4, 0x1.f295c43b61425p+3},
    {-0x1.61d41b2d54580p+0, 0x1.497af24bb677ep+3, -0x1.1d406ee60becfp+3},
    {0x1.5555555555555p-2, -0x1.5555555555555p+1, 0x1.aaaaaaaaaaaabp+1}};

/* ---- options: field-for-field the constructor keywords of radauDAE.py:263-278 ---- */
typedef struct {
    double first_step;  /* <=0: reference default with a mass matrix, min(1e-6,|tb-t0|) (:313) */
    double max_step;
    double newton_tol;  /* <=0: max(10 eps/rtol, min(0.03, sqrt(rtol))) (:322) */
    double factor_on_non_convergence, safety_factor, min_factor, max_factor;
    double jac_recompute_factor;
    int32_t max_newton_ite;
    int32_t max_bad_ite; /* <0: 1 if any var_index>1 else 0 (:406-410) */
    int32_t scale_residuals, scale_newton_norm, scale_error, zero_algebraic_error;
    int32_t always_2nd_estimate, predictive_controller, predictive_newton_stop;
    int32_t extrapolated_guess, all_newton_iterations, constant_dt;
    int32_t jac_finite_differences; /* jac=None: scipy num_jac forward differences (:468-481) */
    int32_t reserved;
    int64_t nmax_step;  /* <=0: unlimited (solve_ivp_custom :1081) */
} orc_options;

enum { ST_FINISHED = 0, ST_TOO_SMALL = -1, ST_LU_FAILED = -2, ST_NMAX = 2 };
enum { C_NFEV, C_NJEV, C_NLU, C_NSTEP, C_NACC, C_NREJ, C_NFAIL, C_NSOLVE, C_NSOLVE_ERR, C_COUNT };

/* ---- problems: same parameter vectors as dae_scipy_b200/problems.py ---- */
enum { PB_PENDULUM3 = 0, PB_PENDULUM2 = 1, PB_PENDULUM1 = 2, PB_ROBERTSON = 3, PB_TRANSISTOR = 4,
       PB_NPENDULUM = 5, PB_ROBERTSON_ODE = 6, PB_JAY = 7, PB_POLYDAE2 = 8, PB_LINEAR7 = 9 };

typedef struct {
    int problem, n;
    const double *p; /* this system's parameters */
    /* n-pendulum derived constants */
    int nn;
    double *m, *inv_m, *r0s;
} prob_ctx;

/* pendulum.py:130-148 (index 3), :151-166 (index 2), :169-184 (index 1); p = (m, g, r0) */
static void pend_rhs(const prob_ctx *c, double t, const double *X, double *f) {
    (void)t;
    /* contract: multiply by im = 1/m where the reference divides by m (identical for m = 1) */
    double im = 1.0 / c->p[0], g = c->p[1], r0 = c->p[2];
    double x = X[0], y = X[1], vx = X[2], vy = X[3], l = X[4];
    f[0] = vx;
    f[1] = vy;
    f[2] = -x * l * im;
    f[3] = -g - (y * l) * im;
    if (c->problem == PB_PENDULUM3) f[4] = x * x + y * y - r0 * r0;
    else if (c->problem == PB_PENDULUM2) f[4] = x * vx + y * vy;
    else f[4] = l * (x * x + y * y) * im + g * y - (vx * vx + vy * vy);
}
static void pend_jac(const prob_ctx *c, double t, const double *X, double *J) {
    (void)t;
    double im = 1.0 / c->p[0], g = c->p[1];
    double x = X[0], y = X[1], vx = X[2], vy = X[3], l = X[4];
    memset(J, 0, 25 * sizeof(double));
    J[0 * 5 + 2] = 1.0;
    J[1 * 5 + 3] = 1.0;
    J[2 * 5 + 0] = -l * im; J[2 * 5 + 4] = -x * im;
    J[3 * 5 + 1] = -l * im; J[3 * 5 + 4] = -y * im;
    if (c->problem == PB_PENDULUM3) { J[4 * 5 + 0] = 2 * x; J[4 * 5 + 1] = 2 * y; }
    else if (c->problem == PB_PENDULUM2) { J[20] = vx; J[21] = vy; J[22] = x; J[23] = y; }
    else {
        J[20] = 2 * x * l * im; J[21] = 2 * y * l * im + g; J[22] = -2 * vx; J[23] = -2 * vy;
        J[24] = (x * x + y * y) * im;
    }
}
/* robertson.py:80-87; p = (k1,k2,k3) */
static void rob_rhs(const prob_ctx *c, double t, const double *y, double *f) {
    (void)t;
    double k1 = c->p[0], k2 = c->p[1], k3 = c->p[2];
    f[0] = -k1 * y[0] + k2 * y[1] * y[2];
    f[1] = k1 * y[0] - k2 * y[1] * y[2] - k3 * (y[1] * y[1]);
    f[2] = y[0] + y[1] + y[2] - 1;
}
static void rob_jac(const prob_ctx *c, double t, const double *y, double *J) {
    (void)t;
    double k1 = c->p[0], k2 = c->p[1], k3 = c->p[2];
    J[0] = -k1; J[1] = k2 * y[2]; J[2] = k2 * y[1];
    J[3] = k1;  J[4] = -k2 * y[2] - 2 * k3 * y[1]; J[5] = -k2 * y[1];
    J[6] = 1; J[7] = 1; J[8] = 1;
}
/* robertson.py:69-77 (ODE formulation) */
static void robode_rhs(const prob_ctx *c, double t, const double *y, double *f) {
    (void)t;
    double k1 = c->p[0], k2 = c->p[1], k3 = c->p[2];
    f[0] = -k1 * y[0] + k2 * y[1] * y[2];
    f[1] = k1 * y[0] - k2 * y[1] * y[2] - k3 * (y[1] * y[1]);
    f[2] = k3 * (y[1] * y[1]);
}
static void robode_jac(const prob_ctx *c, double t, const double *y, double *J) {
    (void)t;
    double k1 = c->p[0], k2 = c->p[1], k3 = c->p[2];
    J[0] = -k1; J[1] = k2 * y[2]; J[2] = k2 * y[1];
    J[3] = k1;  J[4] = -k2 * y[2] - 2 * k3 * y[1]; J[5] = -k2 * y[1];
    J[6] = 0; J[7] = 2 * k3 * y[1]; J[8] = 0;
}
/* jay_dae.py:22-45; p = (nonlinear flag) */
static void jay_rhs(const prob_ctx *c, double t, const double *y, double *f) {
    (void)t;
    double y1 = y[0], y2 = y[1], z1 = y[2], z2 = y[3], la = y[4];
    f[0] = 2 * y1 * y2 * z1 * z2;
    f[1] = -y1 * y2 * (z2 * z2);
    f[2] = (y1 * y2 + z1 * z2) * la;
    if (c->p[0] != 0.0) f[3] = -y1 * (y2 * y2) * (z2 * z2 * z2) * (la * la);
    else f[3] = -y1 * (y2 * y2) * (z2 * z2) * la;
    f[4] = y1 * (y2 * y2) - 1.0;
}
static void jay_jac(const prob_ctx *c, double t, const double *y, double *J) {
    (void)t;
    double y1 = y[0], y2 = y[1], z1 = y[2], z2 = y[3], la = y[4];
    memset(J, 0, 25 * sizeof(double));
    J[0] = 2 * y2 * z1 * z2; J[1] = 2 * y1 * z1 * z2; J[2] = 2 * y1 * y2 * z2; J[3] = 2 * y1 * y2 * z1;
    J[5] = -y2 * (z2 * z2); J[6] = -y1 * (z2 * z2); J[8] = -2 * y1 * y2 * z2;
    J[10] = y2 * la; J[11] = y1 * la; J[12] = z2 * la; J[13] = z1 * la; J[14] = y1 * y2 + z1 * z2;
    if (c->p[0] != 0.0) {
        J[15] = -(y2 * y2) * (z2 * z2 * z2) * (la * la); J[16] = -2 * y1 * y2 * (z2 * z2 * z2) * (la * la);
        J[18] = -3 * y1 * (y2 * y2) * (z2 * z2) * (la * la); J[19] = -2 * y1 * (y2 * y2) * (z2 * z2 * z2) * la;
    } else {
        J[15] = -(y2 * y2) * (z2 * z2) * la; J[16] = -2 * y1 * y2 * (z2 * z2) * la;
        J[18] = -2 * y1 * (y2 * y2) * z2 * la; J[19] = -y1 * (y2 * y2) * (z2 * z2);
    }
    J[20] = y2 * y2; J[21] = 2 * y1 * y2;
}
/* validation_dense_output.py:23-59, p = 2; params (a4,a3,a2,a1,a0), Horner evaluation */
static void poly_rhs(const prob_ctx *c, double t, const double *x, double *f) {
    const double *a = c->p;
    double P = (((a[0] * t + a[1]) * t + a[2]) * t + a[3]) * t + a[4];
    double dP = ((4 * a[0] * t + 3 * a[1]) * t + 2 * a[2]) * t + a[3];
    f[0] = dP + (x[1] - x[0]);
    f[1] = x[0] - P;
}
static void poly_jac(const prob_ctx *c, double t, const double *x, double *J) {
    (void)c; (void)t; (void)x;
    J[0] = -1; J[1] = 1; J[2] = 1; J[3] = 0;
}
/* linear_ODE_with_mass_matrix.py:32-39: f = c * x, params c[7] */
static void lin_rhs(const prob_ctx *c, double t, const double *x, double *f) {
    (void)t;
    for (int k = 0; k < 7; ++k) f[k] = c->p[k] * x[k];
}
static void lin_jac(const prob_ctx *c, double t, const double *x, double *J) {
    (void)t; (void)x;
    memset(J, 0, 49 * sizeof(double));
    for (int k = 0; k < 7; ++k) J[k * 7 + k] = c->p[k];
}
/* transistor_amplifier.py:33-56; p = (C1,C2,C3,R0,R1,R2,R3,R4,R5,Ub,Ua,w) */
static void tra_rhs(const prob_ctx *c, double t, const double *y, double *f) {
    const double *p = c->p;
    /* contract: conductances 1/R formed once; multiply where the reference divides */
    double R0 = p[3], R1 = p[4], R2 = p[5], R3 = p[6], R4 = p[7], R5 = p[8], Ub = p[9];
    double g0 = 1 / R0, g12 = 1 / R1 + 1 / R2, g3 = 1 / R3, g4 = 1 / R4, g5 = 1 / R5, ub2 = Ub / R2, ub4 = Ub / R4;
    double Ue = p[10] * sin(p[11] * t);
    double fT = 1e-6 * (exp((y[1] - y[2]) * (1 / 0.026)) - 1);
    f[0] = (Ue - y[0]) * g0;
    f[1] = ub2 - y[1] * g12 - 0.01 * fT;
    f[2] = fT - y[2] * g3;
    f[3] = ub4 - y[3] * g4 - 0.99 * fT;
    f[4] = -y[4] * g5;
}
static void tra_jac(const prob_ctx *c, double t, const double *y, double *J) {
    (void)t;
    const double *p = c->p;
    double R0 = p[3], R1 = p[4], R2 = p[5], R3 = p[6], R4 = p[7], R5 = p[8];
    double g0 = 1 / R0, g12 = 1 / R1 + 1 / R2, g3 = 1 / R3, g4 = 1 / R4, g5 = 1 / R5;
    double df = 1e-6 * exp((y[1] - y[2]) * (1 / 0.026)) * (1 / 0.026);
    memset(J, 0, 25 * sizeof(double));
    J[0] = -g0;
    J[6] = -g12 - 0.01 * df; J[7] = 0.01 * df;
    J[11] = df; J[12] = -df - g3;
    J[16] = -0.99 * df; J[17] = 0.99 * df; J[18] = -g4;
    J[24] = -g5;
}
static void tra_mass(const double *p, double *M) {
    double C1 = p[0], C2 = p[1], C3 = p[2];
    memset(M, 0, 25 * sizeof(double));
    M[0] = C1; M[1] = -C1; M[5] = -C1; M[6] = C1; M[12] = C2;
    M[18] = C3; M[19] = -C3; M[23] = -C3; M[24] = C3;
}
/* recursive_pendulum.py:93-118; unknowns interleaved (x,y,vx,vy,lambda) per node */
#define GRAV 9.81
static void npe_rhs(const prob_ctx *c, double t, const double *X, double *f) {
    (void)t;
    int nn = c->nn;
    double a_prev = 0, b_prev = 0; /* lambda_i dx_i / rs_i of node i, consumed by node i-1 */
    /* walk from the last node to the first so that the (i+1) term is at hand */
    for (int i = nn - 1; i >= 0; --i) {
        const double *q = X + 5 * i;
        double dx = (i == 0) ? q[0] : q[0] - q[-5];
        double dy = (i == 0) ? q[1] : q[1] - q[-4];
        double rs = dx * dx + dy * dy;
        double ir = 1.0 / rs;                 /* one reciprocal per rod (chain spec) */
        double a = (q[4] * dx) * ir, b = (q[4] * dy) * ir;
        double *o = f + 5 * i;
        o[0] = q[2];
        o[1] = q[3];
        if (i == nn - 1) {
            o[2] = c->inv_m[i] * a;
            o[3] = c->inv_m[i] * (b - c->m[i] * GRAV);
        } else {
            o[2] = c->inv_m[i] * (a - a_prev);
            o[3] = c->inv_m[i] * (b - b_prev - c->m[i] * GRAV);
        }
        o[4] = rs - c->r0s[i];
        a_prev = a; b_prev = b;
    }
}
/* analytic Jacobian of npe_rhs, dense [N][N]; same formulas as oracle/problems_numpy.py */
static void npe_jac(const prob_ctx *c, double t, const double *X, double *J) {
    (void)t;
    int nn = c->nn, N = 5 * nn;
    memset(J, 0, (size_t)N * N * sizeof(double));
    for (int i = 0; i < nn; ++i) {
        const double *q = X + 5 * i;
        double dx = (i == 0) ? q[0] : q[0] - q[-5];
        double dy = (i == 0) ? q[1] : q[1] - q[-4];
        double rs = dx * dx + dy * dy, l = q[4];
        double rs2 = rs * rs;
        double adx = l * (rs - 2 * dx * dx) / rs2, ady = l * (-2 * dx * dy) / rs2;
        double bdy = l * (rs - 2 * dy * dy) / rs2, bdx = ady;
        double al = dx / rs, bl = dy / rs;
        int r = 5 * i;
        double im = c->inv_m[i];
#define JJ(a, b) J[(size_t)(a) * N + (b)]
        JJ(r, r + 2) = 1.0;
        JJ(r + 1, r + 3) = 1.0;
        JJ(r + 2, r) += im * adx; JJ(r + 2, r + 1) += im * ady; JJ(r + 2, r + 4) += im * al;
        JJ(r + 3, r) += im * bdx; JJ(r + 3, r + 1) += im * bdy; JJ(r + 3, r + 4) += im * bl;
        if (i > 0) {
            JJ(r + 2, r - 5) -= im * adx; JJ(r + 2, r - 4) -= im * ady;
            JJ(r + 3, r - 5) -= im * bdx; JJ(r + 3, r - 4) -= im * bdy;
            /* node i-1 sees -a_i, -b_i */
            double imp = c->inv_m[i - 1];
            int rp = r - 5;
            JJ(rp + 2, r) -= imp * adx; JJ(rp + 2, r + 1) -= imp * ady; JJ(rp + 2, r + 4) -= imp * al;
            JJ(rp + 2, rp) += imp * adx; JJ(rp + 2, rp + 1) += imp * ady;
            JJ(rp + 3, r) -= imp * bdx; JJ(rp + 3, r + 1) -= imp * bdy; JJ(rp + 3, r + 4) -= imp * bl;
            JJ(rp + 3, rp) += imp * bdx; JJ(rp + 3, rp + 1) += imp * bdy;
        }
        JJ(r + 4, r) = 2 * dx; JJ(r + 4, r + 1) = 2 * dy;
        if (i > 0) { JJ(r + 4, r - 5) = -2 * dx; JJ(r + 4, r - 4) = -2 * dy; }
#undef JJ
    }
}

/* recursive_pendulum.py:63-73: node masses 1/(n-1) (last one 1), rods 2/n */
static void npe_constants(prob_ctx *c) {
    int nn = c->nn;
    c->m = malloc(sizeof(double) * nn); c->inv_m = malloc(sizeof(double) * nn); c->r0s = malloc(sizeof(double) * nn);
    for (int i = 0; i < nn; ++i) {
        double m = (i == nn - 1) ? 1.0 : 1.0 / (nn - 1);
        double r0 = 2.0 / nn;
        c->m[i] = m; c->inv_m[i] = 1 / m; c->r0s[i] = r0 * r0;
    }
}

static void prob_rhs(const prob_ctx *c, double t, const double *y, double *f) {
    switch (c->problem) {
    case PB_ROBERTSON: rob_rhs(c, t, y, f); break;
    case PB_TRANSISTOR: tra_rhs(c, t, y, f); break;
    case PB_NPENDULUM: npe_rhs(c, t, y, f); break;
    case PB_ROBERTSON_ODE: robode_rhs(c, t, y, f); break;
    case PB_JAY: jay_rhs(c, t, y, f); break;
    case PB_POLYDAE2: poly_rhs(c, t, y, f); break;
    case PB_LINEAR7: lin_rhs(c, t, y, f); break;
    default: pend_rhs(c, t, y, f);
    }
}
static void prob_jac(const prob_ctx *c, double t, const double *y, double *J) {
    switch (c->problem) {
    case PB_ROBERTSON: rob_jac(c, t, y, J); break;
    case PB_TRANSISTOR: tra_jac(c, t, y, J); break;
    case PB_NPENDULUM: npe_jac(c, t, y, J); break;
    case PB_ROBERTSON_ODE: robode_jac(c, t, y, J); break;
    case PB_JAY: jay_jac(c, t, y, J); break;
    case PB_POLYDAE2: poly_jac(c, t, y, J); break;
    case PB_LINEAR7: lin_jac(c, t, y, J); break;
    default: pend_jac(c, t, y, J);
    }
}

int orc_problem_dims(int problem, int *n, int *n_params) {
    switch (problem) {
    case PB_PENDULUM3: case PB_PENDULUM2: case PB_PENDULUM1: *n = 5; *n_params = 3; return 0;
    case PB_ROBERTSON: *n = 3; *n_params = 3; return 0;
    case PB_TRANSISTOR: *n = 5; *n_params = 12; return 0;
    case PB_NPENDULUM: *n = -1; *n_params = 0; return 0; /* n = 5*nodes given by the caller */
    case PB_ROBERTSON_ODE: *n = 3; *n_params = 3; return 0;
    case PB_JAY: *n = 5; *n_params = 1; return 0;
    case PB_POLYDAE2: *n = 2; *n_params = 5; return 0;
    case PB_LINEAR7: *n = 7; *n_params = 7; return 0;
    }
    return -1;
}

/* test hooks: evaluate rhs / jacobian of one system (used to check the problem restatements) */
int orc_eval_rhs_jac(int problem, int n, const double *p, double t, const double *y, double *f,
                     double *J) {
    prob_ctx c = {problem, n, p, 0, NULL, NULL, NULL};
    if (problem == PB_NPENDULUM) { c.nn = n / 5; npe_constants(&c); }
    if (f) prob_rhs(&c, t, y, f);
    if (J) prob_jac(&c, t, y, J);
    free(c.m); free(c.inv_m); free(c.r0s);
    return 0;
}

/* ---- linear algebra of the spec ---- */
typedef struct {
    int n, kl, ku;       /* band limits used only to skip structural zeros (results identical) */
    double *lr;          /* real LU [n][n], reciprocal pivots on the diagonal */
    double *lcr, *lci;   /* complex LU, split re/im, reciprocal pivots on the diagonal */
    int *pr, *pc;        /* pivot rows */
} lu_pair;

static int imin(int a, int b) { return a < b ? a : b; }

/* right-looking LU, partial pivoting on |a| (first maximum), reciprocal pivot stored */
static void lu_real(lu_pair *L) {
    int n = L->n;
    double *A = L->lr;
    for (int k = 0; k < n; ++k) {
        int rmax = imin(n - 1, k + L->kl);
        int cmax = imin(n - 1, k + L->ku + L->kl);
        int p = k;
        double best = fabs(A[k * n + k]);
        for (int r = k + 1; r <= rmax; ++r) {
            double v = fabs(A[r * n + k]);
            if (v > best) { best = v; p = r; }
        }
        L->pr[k] = p;
        if (p != k) /* active part only: multipliers stay where they were computed (dgbtrf style) */
            for (int c = k; c <= cmax; ++c) { double tmp = A[k * n + c]; A[k * n + c] = A[p * n + c]; A[p * n + c] = tmp; }
        double rcp = 1.0 / A[k * n + k];
        A[k * n + k] = rcp;
        for (int r = k + 1; r <= rmax; ++r) {
            double l = A[r * n + k] * rcp;
            A[r * n + k] = l;
            for (int c = k + 1; c <= cmax; ++c) A[r * n + c] = fma(-l, A[k * n + c], A[r * n + c]);
        }
    }
}
static void solve_real(const lu_pair *L, double *b) {
    int n = L->n;
    const double *A = L->lr;
    for (int k = 0; k < n; ++k) {
        int p = L->pr[k];
        if (p != k) { double tmp = b[k]; b[k] = b[p]; b[p] = tmp; }
        int rmax = imin(n - 1, k + L->kl);
        double bk = b[k];
        for (int r = k + 1; r <= rmax; ++r) b[r] = fma(-A[r * n + k], bk, b[r]);
    }
    /* back substitution, column oriented: x_k is final once columns > k have been applied */
    for (int k = n - 1; k >= 0; --k) {
        int rmin = k - (L->ku + L->kl); if (rmin < 0) rmin = 0;
        double xk = b[k] * A[k * n + k];
        b[k] = xk;
        for (int r = rmin; r < k; ++r) b[r] = fma(-A[r * n + k], xk, b[r]);
    }
}
/* complex: pivot on |re|+|im| (BLAS izamax), reciprocal pivot = conj(p)/|p|^2 */
static void lu_cplx(lu_pair *L) {
    int n = L->n;
    double *R = L->lcr, *I = L->lci;
    for (int k = 0; k < n; ++k) {
        int rmax = imin(n - 1, k + L->kl);
        int cmax = imin(n - 1, k + L->ku + L->kl);
        int p = k;
        double best = fabs(R[k * n + k]) + fabs(I[k * n + k]);
        for (int r = k + 1; r <= rmax; ++r) {
            double v = fabs(R[r * n + k]) + fabs(I[r * n + k]);
            if (v > best) { best = v; p = r; }
        }
        L->pc[k] = p;
        if (p != k)
            for (int c = k; c <= cmax; ++c) {
                double tmp = R[k * n + c]; R[k * n + c] = R[p * n + c]; R[p * n + c] = tmp;
                tmp = I[k * n + c]; I[k * n + c] = I[p * n + c]; I[p * n + c] = tmp;
            }
        double pr = R[k * n + k], pi = I[k * n + k];
        double invd = 1.0 / fma(pr, pr, pi * pi);
        double qr = pr * invd, qi = -(pi * invd);
        R[k * n + k] = qr; I[k * n + k] = qi;
        for (int r = k + 1; r <= rmax; ++r) {
            double ar = R[r * n + k], ai = I[r * n + k];
            double lr = fma(ar, qr, -(ai * qi));
            double li = fma(ar, qi, ai * qr);
            R[r * n + k] = lr; I[r * n + k] = li;
            for (int c = k + 1; c <= cmax; ++c) {
                double ur = R[k * n + c], ui = I[k * n + c];
                double xr = R[r * n + c], xi = I[r * n + c];
                xr = fma(-lr, ur, xr); xr = fma(li, ui, xr);
                xi = fma(-lr, ui, xi); xi = fma(-li, ur, xi);
                R[r * n + c] = xr; I[r * n + c] = xi;
            }
        }
    }
}
static void solve_cplx(const lu_pair *L, double *br, double *bi) {
    int n = L->n;
    const double *R = L->lcr, *I = L->lci;
    for (int k = 0; k < n; ++k) {
        int p = L->pc[k];
        if (p != k) {
            double tmp = br[k]; br[k] = br[p]; br[p] = tmp;
            tmp = bi[k]; bi[k] = bi[p]; bi[p] = tmp;
        }
        int rmax = imin(n - 1, k + L->kl);
        double xr = br[k], xi = bi[k];
        for (int r = k + 1; r <= rmax; ++r) {
            double lr = R[r * n + k], li = I[r * n + k];
            double yr = br[r], yi = bi[r];
            yr = fma(-lr, xr, yr); yr = fma(li, xi, yr);
            yi = fma(-lr, xi, yi); yi = fma(-li, xr, yi);
            br[r] = yr; bi[r] = yi;
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        int rmin = k - (L->ku + L->kl); if (rmin < 0) rmin = 0;
        double qr = R[k * n + k], qi = I[k * n + k];
        double sr = br[k], si = bi[k];
        double xr = fma(sr, qr, -(si * qi));
        double xi = fma(sr, qi, si * qr);
        br[k] = xr; bi[k] = xi;
        for (int r = rmin; r < k; ++r) {
            double ur = R[r * n + k], ui = I[r * n + k];
            double yr = br[r], yi = bi[r];
            yr = fma(-ur, xr, yr); yr = fma(ui, xi, yr);
            yi = fma(-ur, xi, yi); yi = fma(-ui, xr, yi);
            br[r] = yr; bi[r] = yi;
        }
    }
}

/* ---- chain block elimination: the linear algebra of the n-pendulum ("chain spec") ----
 * The iteration matrix mu M - J of the chain (recursive_pendulum.py:93-127; M = diag(1,1,1,1,0) per node) is
 * block tridiagonal in the nodes.  Rows x_i, y_i read  mu p - v = r, so the velocity increments are eliminated
 * exactly (v = mu p - r) and what is left per node is a 3 x 3 block in u_i = (dx_i, dy_i, dlambda_i):
 *     D_i u_i + L_i u_{i-1} + U_i u_{i+1} = g_i,   g = (r_vx + mu r_x, r_vy + mu r_y, r_lambda),
 * L_i with a zero third column and U_i with a zero third row.  It is factorised node by node,
 *     E_0 = D_0,  E_i = D_i - L_i E_{i-1}^-1 U_{i-1},
 * each E_i inverted explicitly by Gauss-Jordan elimination with partial pivoting inside the block (first
 * maximum of |a|, |re|+|im| for the complex system, reciprocal pivots), and solved by
 *     z_i = E_i^-1 (g_i - L_i z_{i-1}),   u_i = z_i - E_i^-1 U_i u_{i+1}.
 * The matrix is NOT row-scaled by hscale (radauDAE.py:600-606): in the Newton iteration the scaling cancels, and
 * the one place where the reference applies the scaled factors to an unscaled right-hand side -- the error
 * estimate, SURVEY Q9 -- multiplies those rows of the right-hand side by the h of the factorisation instead.
 * Every product/sum below is written in the order the CUDA kernel (csrc/k4_chain.cu) evaluates it. */
typedef struct { double adx, ady, bdy, al, bl, dx2, dy2; } cnode;

static void chain_jac(const prob_ctx *c, const double *X, cnode *nd) {
    for (int i = 0; i < c->nn; ++i) {
        const double *q = X + 5 * i;
        double dx = (i == 0) ? q[0] : q[0] - q[-5];
        double dy = (i == 0) ? q[1] : q[1] - q[-4];
        double rs = dx * dx + dy * dy, l = q[4];
        double rs2 = rs * rs;
        nd[i].adx = l * (rs - 2 * dx * dx) / rs2;
        nd[i].ady = l * (-2 * dx * dy) / rs2;
        nd[i].bdy = l * (rs - 2 * dy * dy) / rs2;
        nd[i].al = dx / rs;
        nd[i].bl = dy / rs;
        nd[i].dx2 = 2 * dx;
        nd[i].dy2 = 2 * dy;
    }
}
/* V = E^-1, Gauss-Jordan with partial pivoting; E is destroyed */
static void inv3_real(double E[3][3], double V[3][3]) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
    for (int k = 0; k < 3; ++k) {
        int p = k;
        double best = fabs(E[k][k]);
        for (int r = k + 1; r < 3; ++r) { double v = fabs(E[r][k]); if (v > best) { best = v; p = r; } }
        if (p != k)
            for (int c = 0; c < 3; ++c) {
                double t = E[k][c]; E[k][c] = E[p][c]; E[p][c] = t;
                t = V[k][c]; V[k][c] = V[p][c]; V[p][c] = t;
            }
        double rcp = 1.0 / E[k][k];
        for (int c = k + 1; c < 3; ++c) E[k][c] = E[k][c] * rcp;
        for (int c = 0; c < 3; ++c) V[k][c] = V[k][c] * rcp;
        for (int r = 0; r < 3; ++r) {
            if (r == k) continue;
            double f = E[r][k];
            for (int c = k + 1; c < 3; ++c) E[r][c] = fma(-f, E[k][c], E[r][c]);
            for (int c = 0; c < 3; ++c) V[r][c] = fma(-f, V[k][c], V[r][c]);
        }
    }
}
static void inv3_cplx(double Er[3][3], double Ei[3][3], double Vr[3][3], double Vi[3][3]) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { Vr[r][c] = (r == c) ? 1.0 : 0.0; Vi[r][c] = 0.0; }
    for (int k = 0; k < 3; ++k) {
        int p = k;
        double best = fabs(Er[k][k]) + fabs(Ei[k][k]);
        for (int r = k + 1; r < 3; ++r) { double v = fabs(Er[r][k]) + fabs(Ei[r][k]); if (v > best) { best = v; p = r; } }
        if (p != k)
            for (int c = 0; c < 3; ++c) {
                double t = Er[k][c]; Er[k][c] = Er[p][c]; Er[p][c] = t;
                t = Ei[k][c]; Ei[k][c] = Ei[p][c]; Ei[p][c] = t;
                t = Vr[k][c]; Vr[k][c] = Vr[p][c]; Vr[p][c] = t;
                t = Vi[k][c]; Vi[k][c] = Vi[p][c]; Vi[p][c] = t;
            }
        double pr = Er[k][k], pi = Ei[k][k];
        double invd = 1.0 / fma(pr, pr, pi * pi);
        double qr = pr * invd, qi = -(pi * invd);
        for (int c = k + 1; c < 3; ++c) {
            double xr = Er[k][c], xi = Ei[k][c];
            Er[k][c] = fma(xr, qr, -(xi * qi)); Ei[k][c] = fma(xr, qi, xi * qr);
        }
        for (int c = 0; c < 3; ++c) {
            double xr = Vr[k][c], xi = Vi[k][c];
            Vr[k][c] = fma(xr, qr, -(xi * qi)); Vi[k][c] = fma(xr, qi, xi * qr);
        }
        for (int r = 0; r < 3; ++r) {
            if (r == k) continue;
            double fr = Er[r][k], fi = Ei[r][k];
            for (int c = k + 1; c < 3; ++c) {
                double ur = Er[k][c], ui = Ei[k][c], xr = Er[r][c], xi = Ei[r][c];
                xr = fma(-fr, ur, xr); xr = fma(fi, ui, xr);
                xi = fma(-fr, ui, xi); xi = fma(-fi, ur, xi);
                Er[r][c] = xr; Ei[r][c] = xi;
            }
            for (int c = 0; c < 3; ++c) {
                double ur = Vr[k][c], ui = Vi[k][c], xr = Vr[r][c], xi = Vi[r][c];
                xr = fma(-fr, ur, xr); xr = fma(fi, ui, xr);
                xi = fma(-fr, ui, xi); xi = fma(-fi, ur, xi);
                Vr[r][c] = xr; Vi[r][c] = xi;
            }
        }
    }
}
typedef struct {
    int nn;
    cnode *nd;                 /* rod quantities of the Jacobian, [nn] */
    double *vr;                /* E_i^-1 of the real system, [nn][9] */
    double *vcr, *vci;         /* E_i^-1 of the complex system */
    double *zr, *zcr, *zci;    /* forward-substituted blocks, [nn][3] */
} chain_lu;

/* coupling blocks of node i: L_i[:, 0:2] (3 x 2, to node i-1) and U_{i-1}[0:2, :] (2 x 3, node i-1 to node i) */
static void chain_L(const prob_ctx *c, const cnode *nd, int i, double L[3][2]) {
    double im = c->inv_m[i];
    L[0][0] = im * nd[i].adx; L[0][1] = im * nd[i].ady;
    L[1][0] = im * nd[i].ady; L[1][1] = im * nd[i].bdy;
    L[2][0] = nd[i].dx2;      L[2][1] = nd[i].dy2;
}
static void chain_U(const prob_ctx *c, const cnode *nd, int i /* block U_i couples node i to node i+1 */, double U[2][3]) {
    double im = c->inv_m[i];
    U[0][0] = im * nd[i + 1].adx; U[0][1] = im * nd[i + 1].ady; U[0][2] = im * nd[i + 1].al;
    U[1][0] = im * nd[i + 1].ady; U[1][1] = im * nd[i + 1].bdy; U[1][2] = im * nd[i + 1].bl;
}
/* both factorisations; mr, (mcr, mci) = mu / h */
static void chain_factor(const prob_ctx *c, chain_lu *C, double mr, double mcr, double mci) {
    const int nn = C->nn;
    const cnode *nd = C->nd;
    const double m2 = mr * mr;
    const double m2r = fma(mcr, mcr, -(mci * mci)), m2i = 2 * (mcr * mci);
    for (int i = 0; i < nn; ++i) {
        double im = c->inv_m[i];
        double sxx = im * nd[i].adx, sxy = im * nd[i].ady, syy = im * nd[i].bdy;
        if (i < nn - 1) { sxx = sxx + im * nd[i + 1].adx; sxy = sxy + im * nd[i + 1].ady; syy = syy + im * nd[i + 1].bdy; }
        double e02 = -(im * nd[i].al), e12 = -(im * nd[i].bl);
        double Er[3][3] = {{m2 - sxx, -sxy, e02}, {-sxy, m2 - syy, e12}, {-nd[i].dx2, -nd[i].dy2, 0.0}};
        double Cr[3][3] = {{m2r - sxx, -sxy, e02}, {-sxy, m2r - syy, e12}, {-nd[i].dx2, -nd[i].dy2, 0.0}};
        double Ci[3][3] = {{m2i, 0.0, 0.0}, {0.0, m2i, 0.0}, {0.0, 0.0, 0.0}};
        if (i > 0) {
            double L[3][2], U[2][3];
            chain_L(c, nd, i, L);
            chain_U(c, nd, i - 1, U);
            const double *G = C->vr + 9 * (i - 1), *Gr = C->vcr + 9 * (i - 1), *Gi = C->vci + 9 * (i - 1);
            double X[2][3], Xr[2][3], Xi[2][3];
            for (int r = 0; r < 2; ++r)
                for (int k = 0; k < 3; ++k) {
                    X[r][k] = fma(G[3 * r + 1], U[1][k], G[3 * r] * U[0][k]);
                    Xr[r][k] = fma(Gr[3 * r + 1], U[1][k], Gr[3 * r] * U[0][k]);
                    Xi[r][k] = fma(Gi[3 * r + 1], U[1][k], Gi[3 * r] * U[0][k]);
                }
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 3; ++k) {
                    Er[r][k] = Er[r][k] - fma(L[r][1], X[1][k], L[r][0] * X[0][k]);
                    Cr[r][k] = Cr[r][k] - fma(L[r][1], Xr[1][k], L[r][0] * Xr[0][k]);
                    Ci[r][k] = Ci[r][k] - fma(L[r][1], Xi[1][k], L[r][0] * Xi[0][k]);
                }
        }
        double V[3][3], Vr[3][3], Vi[3][3];
        inv3_real(Er, V);
        inv3_cplx(Cr, Ci, Vr, Vi);
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) {
                C->vr[9 * i + 3 * r + k] = V[r][k]; C->vcr[9 * i + 3 * r + k] = Vr[r][k]; C->vci[9 * i + 3 * r + k] = Vi[r][k];
            }
    }
}
/* b (5 per node: x, y, vx, vy, lambda) <- (mu M - J)^-1 b, real system */
static void chain_solve_real(const prob_ctx *c, chain_lu *C, double mr, double *b) {
    const int nn = C->nn;
    const cnode *nd = C->nd;
    double *z = C->zr;
    for (int i = 0; i < nn; ++i) {
        const double *r = b + 5 * i, *V = C->vr + 9 * i;
        double g[3] = {fma(mr, r[0], r[2]), fma(mr, r[1], r[3]), r[4]};
        if (i > 0) {
            double L[3][2];
            chain_L(c, nd, i, L);
            for (int q = 0; q < 3; ++q) g[q] = fma(-L[q][1], z[3 * (i - 1) + 1], fma(-L[q][0], z[3 * (i - 1)], g[q]));
        }
        for (int q = 0; q < 3; ++q) z[3 * i + q] = fma(V[3 * q + 2], g[2], fma(V[3 * q + 1], g[1], V[3 * q] * g[0]));
    }
    double u[3] = {0, 0, 0};
    for (int i = nn - 1; i >= 0; --i) {
        const double *V = C->vr + 9 * i;
        double un[3];
        if (i == nn - 1) { for (int q = 0; q < 3; ++q) un[q] = z[3 * i + q]; }
        else {
            double U[2][3];
            chain_U(c, nd, i, U);
            double w0 = fma(U[0][2], u[2], fma(U[0][1], u[1], U[0][0] * u[0]));
            double w1 = fma(U[1][2], u[2], fma(U[1][1], u[1], U[1][0] * u[0]));
            for (int q = 0; q < 3; ++q) un[q] = fma(-V[3 * q + 1], w1, fma(-V[3 * q], w0, z[3 * i + q]));
        }
        double *r = b + 5 * i;
        double v1 = fma(mr, un[0], -r[0]), v2 = fma(mr, un[1], -r[1]);
        r[0] = un[0]; r[1] = un[1]; r[2] = v1; r[3] = v2; r[4] = un[2];
        u[0] = un[0]; u[1] = un[1]; u[2] = un[2];
    }
}
static void chain_solve_cplx(const prob_ctx *c, chain_lu *C, double mcr, double mci, double *br, double *bi) {
    const int nn = C->nn;
    const cnode *nd = C->nd;
    double *zr = C->zcr, *zi = C->zci;
    for (int i = 0; i < nn; ++i) {
        const double *rr = br + 5 * i, *ri = bi + 5 * i, *Vr = C->vcr + 9 * i, *Vi = C->vci + 9 * i;
        /* g = r_v + mu r_p (complex product) */
        double gr[3] = {fma(-mci, ri[0], fma(mcr, rr[0], rr[2])), fma(-mci, ri[1], fma(mcr, rr[1], rr[3])), rr[4]};
        double gi[3] = {fma(mci, rr[0], fma(mcr, ri[0], ri[2])), fma(mci, rr[1], fma(mcr, ri[1], ri[3])), ri[4]};
        if (i > 0) {
            double L[3][2];
            chain_L(c, nd, i, L);
            for (int q = 0; q < 3; ++q) {
                gr[q] = fma(-L[q][1], zr[3 * (i - 1) + 1], fma(-L[q][0], zr[3 * (i - 1)], gr[q]));
                gi[q] = fma(-L[q][1], zi[3 * (i - 1) + 1], fma(-L[q][0], zi[3 * (i - 1)], gi[q]));
            }
        }
        for (int q = 0; q < 3; ++q) {
            double sr = fma(Vr[3 * q], gr[0], -(Vi[3 * q] * gi[0]));
            double si = fma(Vr[3 * q], gi[0], Vi[3 * q] * gr[0]);
            for (int k = 1; k < 3; ++k) {
                sr = fma(Vr[3 * q + k], gr[k], sr); sr = fma(-Vi[3 * q + k], gi[k], sr);
                si = fma(Vr[3 * q + k], gi[k], si); si = fma(Vi[3 * q + k], gr[k], si);
            }
            zr[3 * i + q] = sr; zi[3 * i + q] = si;
        }
    }
    double ur[3] = {0, 0, 0}, ui[3] = {0, 0, 0};
    for (int i = nn - 1; i >= 0; --i) {
        const double *Vr = C->vcr + 9 * i, *Vi = C->vci + 9 * i;
        double nr[3], ni[3];
        if (i == nn - 1) { for (int q = 0; q < 3; ++q) { nr[q] = zr[3 * i + q]; ni[q] = zi[3 * i + q]; } }
        else {
            double U[2][3];
            chain_U(c, nd, i, U);
            double w0r = fma(U[0][2], ur[2], fma(U[0][1], ur[1], U[0][0] * ur[0]));
            double w0i = fma(U[0][2], ui[2], fma(U[0][1], ui[1], U[0][0] * ui[0]));
            double w1r = fma(U[1][2], ur[2], fma(U[1][1], ur[1], U[1][0] * ur[0]));
            double w1i = fma(U[1][2], ui[2], fma(U[1][1], ui[1], U[1][0] * ui[0]));
            for (int q = 0; q < 3; ++q) {
                double xr = zr[3 * i + q], xi = zi[3 * i + q];
                xr = fma(-Vr[3 * q], w0r, xr); xr = fma(Vi[3 * q], w0i, xr);
                xi = fma(-Vr[3 * q], w0i, xi); xi = fma(-Vi[3 * q], w0r, xi);
                xr = fma(-Vr[3 * q + 1], w1r, xr); xr = fma(Vi[3 * q + 1], w1i, xr);
                xi = fma(-Vr[3 * q + 1], w1i, xi); xi = fma(-Vi[3 * q + 1], w1r, xi);
                nr[q] = xr; ni[q] = xi;
            }
        }
        double *rr = br + 5 * i, *ri = bi + 5 * i;
        /* v = mu p - r */
        double v1r = fma(-mci, ni[0], fma(mcr, nr[0], -rr[0])), v1i = fma(mci, nr[0], fma(mcr, ni[0], -ri[0]));
        double v2r = fma(-mci, ni[1], fma(mcr, nr[1], -rr[1])), v2i = fma(mci, nr[1], fma(mcr, ni[1], -ri[1]));
        rr[0] = nr[0]; rr[1] = nr[1]; rr[2] = v1r; rr[3] = v2r; rr[4] = nr[2];
        ri[0] = ni[0]; ri[1] = ni[1]; ri[2] = v1i; ri[3] = v2i; ri[4] = ni[2];
        for (int q = 0; q < 3; ++q) { ur[q] = nr[q]; ui[q] = ni[q]; }
    }
}
/* chain norms: five fma chains, one per unknown of a node, walked from the last node to the first and over the
 * stacked vectors (m of them) in order; folded as ((s0+s1)+(s2+s3))+s4 */
static double chain_sum_squares(const double *v, int nn, int m) {
    double s[5] = {0, 0, 0, 0, 0};
    for (int i = nn - 1; i >= 0; --i)
        for (int c = 0; c < 5; ++c)
            for (int q = 0; q < m; ++q) { double x = v[(size_t)q * 5 * nn + 5 * i + c]; s[c] = fma(x, x, s[c]); }
    return ((s[0] + s[1]) + (s[2] + s[3])) + s[4];
}

/* M.v with M dense row-major; the fma chain runs over the columns in increasing order and
 * skips structural zeros of M (adding an exact zero changes nothing) */
static void mass_mv(int n, const double *M, const int *mlo, const int *mhi, const double *v, double *o) {
    for (int r = 0; r < n; ++r) {
        double s = 0.0;
        int first = 1;
        for (int c = mlo[r]; c <= mhi[r]; ++c) {
            double m = M[r * n + c];
            if (m == 0.0) continue;
            if (first) { s = m * v[c]; first = 0; }
            else s = fma(m, v[c], s);
        }
        o[r] = s;
    }
}

/* sum of squares of v[0..m-1].  Small systems (thread-per-system kernels, n <= 8): one fma chain in
 * index order.  Large systems (CTA-per-system kernel): 32 interleaved partial chains (entry e goes
 * to chain e mod 32, in increasing e) folded by the binary tree 16, 8, 4, 2, 1 -- the order a warp
 * reduction produces. */
static double sum_squares(const double *v, int m, int lanes32) {
    if (!lanes32) {
        double s = 0;
        for (int e = 0; e < m; ++e) s = fma(v[e], v[e], s);
        return s;
    }
    double p[32];
    for (int j = 0; j < 32; ++j) p[j] = 0.0;
    for (int e = 0; e < m; ++e) p[e & 31] = fma(v[e], v[e], p[e & 31]);
    for (int off = 16; off >= 1; off >>= 1)
        for (int j = 0; j < off; ++j) p[j] = p[j] + p[j + off];
    return p[0];
}

static double hpow(double h, int e) { return e == 0 ? 1.0 : (e == 1 ? h : h * h); }

/* predict_factor, radauDAE.py:52-98 */
static double gustafsson(double h_abs, double h_abs_old, double err, double err_old, int hist, int predictive) {
    double mult;
    if (!predictive || !hist || err == 0.0) mult = 1.0;
    else mult = h_abs / h_abs_old * sqrt(sqrt(err_old / err));
    double base = 1.0 / sqrt(sqrt(err)); /* err == 0 -> +inf, clamped by the caller */
    return (mult < 1.0 ? mult : 1.0) * base;
}

typedef struct {
    int n;
    double *y, *f, *yold, *Q /*[n][3]*/, *W /*[3][n]*/, *Z /*[3][n]*/, *F, *J, *M;
    double *fr, *fcr, *fci, *mw0, *mw1, *mw2, *inv_ns, *inv_es, *ze, *err, *tmp, *hs, *sq;
    double *jfac, *yp, *fn, *fn2; /* finite-difference Jacobian: column factors, perturbed state, f there */
    int *mlo, *mhi;
    lu_pair lu;
    chain_lu ch;     /* n-pendulum: block elimination (allocated when n % 5 == 0) */
} work;

static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s); if (!p) abort(); return p; }

static void work_alloc(work *w, int n) {
    w->n = n;
    size_t N = (size_t)n;
    w->y = xcalloc(N, 8); w->f = xcalloc(N, 8); w->yold = xcalloc(N, 8);
    w->Q = xcalloc(3 * N, 8); w->W = xcalloc(3 * N, 8); w->Z = xcalloc(3 * N, 8); w->F = xcalloc(N, 8);
    w->J = xcalloc(N * N, 8); w->M = xcalloc(N * N, 8);
    w->fr = xcalloc(N, 8); w->fcr = xcalloc(N, 8); w->fci = xcalloc(N, 8);
    w->mw0 = xcalloc(N, 8); w->mw1 = xcalloc(N, 8); w->mw2 = xcalloc(N, 8);
    w->inv_ns = xcalloc(N, 8); w->inv_es = xcalloc(N, 8); w->ze = xcalloc(N, 8);
    w->err = xcalloc(N, 8); w->tmp = xcalloc(N, 8); w->hs = xcalloc(N, 8); w->sq = xcalloc(3 * N, 8);
    w->mlo = xcalloc(N, sizeof(int)); w->mhi = xcalloc(N, sizeof(int));
    w->jfac = xcalloc(N, 8); w->yp = xcalloc(N, 8); w->fn = xcalloc(N, 8); w->fn2 = xcalloc(N, 8);
    w->lu.n = n;
    w->lu.lr = xcalloc(N * N, 8); w->lu.lcr = xcalloc(N * N, 8); w->lu.lci = xcalloc(N * N, 8);
    w->lu.pr = xcalloc(N, sizeof(int)); w->lu.pc = xcalloc(N, sizeof(int));
    int nn = n / 5;
    w->ch.nn = nn;
    w->ch.nd = xcalloc((size_t)nn, sizeof(cnode));
    w->ch.vr = xcalloc(9 * (size_t)nn, 8); w->ch.vcr = xcalloc(9 * (size_t)nn, 8); w->ch.vci = xcalloc(9 * (size_t)nn, 8);
    w->ch.zr = xcalloc(3 * (size_t)nn, 8); w->ch.zcr = xcalloc(3 * (size_t)nn, 8); w->ch.zci = xcalloc(3 * (size_t)nn, 8);
}
static void work_free(work *w) {
    free(w->y); free(w->f); free(w->yold); free(w->Q); free(w->W); free(w->Z); free(w->F);
    free(w->J); free(w->M); free(w->fr); free(w->fcr); free(w->fci); free(w->mw0); free(w->mw1);
    free(w->mw2); free(w->inv_ns); free(w->inv_es); free(w->ze); free(w->err); free(w->tmp);
    free(w->hs); free(w->sq); free(w->mlo); free(w->mhi);
    free(w->jfac); free(w->yp); free(w->fn); free(w->fn2);
    free(w->lu.lr); free(w->lu.lcr); free(w->lu.lci); free(w->lu.pr); free(w->lu.pc);
    free(w->ch.nd); free(w->ch.vr); free(w->ch.vcr); free(w->ch.vci); free(w->ch.zr); free(w->ch.zcr); free(w->ch.zci);
}

/* Z_i = (T W)_i, fma chain over the stages in order 0,1,2; T[2] = (1,1,0) */
static void stage_Z(int n, const double *W, int i, double *z) {
    if (i == 2) { for (int c = 0; c < n; ++c) z[c] = W[c] + W[n + c]; return; }
    for (int c = 0; c < n; ++c) {
        double v = RT[i][0] * W[c];
        v = fma(RT[i][1], W[n + c], v);
        z[c] = fma(RT[i][2], W[2 * n + c], v);
    }
}

/* Optional event trace of ONE system (single-threaded use only): 'F' factorisation pair, 'N' Newton
 * iteration, 'E' error estimate + accept/reject, one char per event.  Used by the scheduling
 * study in profiles/tools/sched_sim.py. */
static char *g_trace = NULL;
static int g_trace_cap = 0, g_trace_len = 0;
static void trace_ev(char c) { if (g_trace && g_trace_len < g_trace_cap) g_trace[g_trace_len++] = c; }

/* Finite-difference Jacobian: scipy.integrate._ivp.common.num_jac (scipy 1.18.1 common.py:268-452,
 * called from radauDAE.py:475-480 with threshold = atol and the column factors kept between calls),
 * dense branch, restated column by column (the reference evaluates all columns in one vectorised
 * call; every operation is element-wise, so the values are the same).  Spec deviation: the columns
 * are scaled by the reciprocal of h instead of divided by h.  Function evaluations made here are
 * not counted in nfev (scipy base.py:143-149 `fun_vectorized`). */
#define NJ_FACTOR0 0x1.0000000000000p-26         /* EPS ** 0.5 */
#define NJ_DIFF_REJECT 0x1.6a09e667f3bcdp-46     /* EPS ** 0.875 */
#define NJ_DIFF_SMALL 0x1.0000000000000p-39      /* EPS ** 0.75 */
#define NJ_DIFF_BIG 0x1.0000000000000p-13        /* EPS ** 0.25 */
#define NJ_MIN_FACTOR 0x1.f400000000000p-43      /* 1e3 * EPS */
static void fd_column(const prob_ctx *pc, int n, double t, const double *y, const double *f, int i, double h,
                      double *yp, double *fn, int *mi_out, double *md_out) {
    for (int c = 0; c < n; ++c) yp[c] = y[c] + (c == i ? h : 0.0);
    prob_rhs(pc, t, yp, fn);
    int mi = 0;
    double md = fabs(fn[0] - f[0]);
    for (int r = 1; r < n; ++r) { const double d = fabs(fn[r] - f[r]); if (d > md) { md = d; mi = r; } }
    *mi_out = mi; *md_out = md;
}
static void num_jac_fd(const prob_ctx *pc, work *w, double t, const double *atol) {
    const int n = w->n;
    const double *y = w->y, *f = w->f;
    for (int i = 0; i < n; ++i) {
        const double ys = (f[i] >= 0 ? 1.0 : -1.0) * fmax(atol[i], fabs(y[i]));
        double fac = w->jfac[i];
        double h = (y[i] + fac * ys) - y[i];
        while (h == 0) { fac *= 10; h = (y[i] + fac * ys) - y[i]; }
        int mi; double md;
        fd_column(pc, n, t, y, f, i, h, w->yp, w->fn, &mi, &md);
        double scale = fmax(fabs(f[mi]), fabs(w->fn[mi]));
        const double *fcol = w->fn;
        if (md < NJ_DIFF_REJECT * scale) {
            const double fac_new = 10 * fac;
            const double h_new = (y[i] + fac_new * ys) - y[i];
            int mi2; double md2;
            fd_column(pc, n, t, y, f, i, h_new, w->yp, w->fn2, &mi2, &md2);
            const double scale2 = fmax(fabs(f[mi2]), fabs(w->fn2[mi2]));
            if (md * scale2 < md2 * scale) { fac = fac_new; h = h_new; fcol = w->fn2; scale = scale2; md = md2; }
        }
        const double rh = 1.0 / h;
        for (int r = 0; r < n; ++r) w->J[r * n + i] = (fcol[r] - f[r]) * rh;
        if (md < NJ_DIFF_SMALL * scale) fac *= 10;
        if (md > NJ_DIFF_BIG * scale) fac *= 0.1;
        w->jfac[i] = fmax(fac, NJ_MIN_FACTOR);
    }
}

/* Per-attempt report, the reference's `_report` (radauDAE.py:424-443, codes :24-28; bReport=True): one record per
 * factorisation (-10), Jacobian update after a failed Newton with a stale Jacobian (-11), non-converged attempt (5), rejected
 * step (2) and accepted step (0), with t, h, the Newton iteration counts and the error norms where the reference passes them.
 * Layouts: count [B]; code [cap][B]; it [cap][2][B] (newton iterations, bad iterations; -1 where not reported);
 * val [cap][4][B] (t, h, err1, err2; err = -1 where not reported). */
typedef struct { int cap; int32_t *count, *code, *it; double *val; } report_ctx;
enum { RC_FACTORISATION = -10, RC_JACOBIAN_UPDATE = -11, RC_NON_CONV = 5, RC_ACCEPTED = 0, RC_REJECTED = 2 };
static void report(const report_ctx *rc, int64_t B, int64_t sys, int code, double t, double h, int newt, int nbad, double e1, double e2) {
    if (!rc || rc->cap <= 0) return;
    int k = rc->count[sys]++;
    if (k >= rc->cap) return;
    rc->code[(int64_t)k * B + sys] = code;
    rc->it[((int64_t)k * 2 + 0) * B + sys] = newt; rc->it[((int64_t)k * 2 + 1) * B + sys] = nbad;
    rc->val[((int64_t)k * 4 + 0) * B + sys] = t; rc->val[((int64_t)k * 4 + 1) * B + sys] = h;
    rc->val[((int64_t)k * 4 + 2) * B + sys] = e1; rc->val[((int64_t)k * 4 + 3) * B + sys] = e2;
}

/* Integrate one system.  Layout of the batch arrays is SoA: element (row, sys) at row*B+sys. */
static void integrate_one(const prob_ctx *pc, work *w, const orc_options *o, const double *mass_in,
                          const double *jac_const, const int *var_index, const double *atol,
                          double rtol, double t0, double tb, const double *t_eval, int T, int64_t B,
                          int64_t sys, const double *y0, double *y_eval, int *n_eval_out,
                          double *t_final, double *y_final, int *status_out, int *counters, const report_ctx *rc) {
    const int n = w->n;
    if (rc && rc->cap > 0) rc->count[sys] = 0;
    int cnt[C_COUNT] = {0};
    const int MAXIT = o->max_newton_ite;
    const double dir = (tb != t0) ? (tb > t0 ? 1.0 : -1.0) : 1.0;
    const double tol = (o->newton_tol > 0) ? o->newton_tol
                       : fmax(10 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));
    int nmax_bad = o->max_bad_ite;
    int n_alg = 0;
    for (int c = 0; c < n; ++c) if (var_index[c] != 0) ++n_alg;
    if (nmax_bad < 0) {
        nmax_bad = 0;
        for (int c = 0; c < n; ++c) if (var_index[c] > 1) nmax_bad = 1;
    }
    const double rsq3n = 1.0 / sqrt(3.0 * n), rsqn = 1.0 / sqrt((double)n);
    const double rsqdiff = 1.0 / sqrt((double)(n - n_alg));
    const int has_jac_fn = (jac_const == NULL);
    /* the chain with its own mass matrix and analytic Jacobian: block elimination (the "chain spec"); reserved = 1
     * keeps the banded LU of the general contract (what the band kernel K2 implements) */
    const int chain = (pc->problem == PB_NPENDULUM && mass_in == NULL && jac_const == NULL
                       && !o->jac_finite_differences && o->reserved == 0);
    double h_lu = 1.0;   /* chain: the h of the current factorisation (rows scaled by hscale in the reference) */
    double mr_f = 0, mcr_f = 0, mci_f = 0;   /* chain: mu / h of the current factorisation (the velocity elimination uses it) */

    /* mass matrix: given (shared) or the problem's own */
    if (mass_in) memcpy(w->M, mass_in, sizeof(double) * n * n);
    else if (pc->problem == PB_TRANSISTOR) tra_mass(pc->p, w->M);
    else { /* pendulum.py:139-140, robertson.py:78, recursive_pendulum.py:120-127: identity with a
              zero on the rows of the algebraic unknowns (last of 5 / last of 3) */
        memset(w->M, 0, sizeof(double) * n * n);
        int blk = (pc->problem == PB_ROBERTSON) ? 3 : (pc->problem == PB_POLYDAE2) ? 2 : 5;
        int identity = (pc->problem == PB_ROBERTSON_ODE || pc->problem == PB_LINEAR7);
        for (int c = 0; c < n; ++c) w->M[c * n + c] = (!identity && c % blk == blk - 1) ? 0.0 : 1.0;
    }
    for (int r = 0; r < n; ++r) {
        int lo = n, hi = -1;
        for (int c = 0; c < n; ++c) if (w->M[r * n + c] != 0.0) { if (c < lo) lo = c; if (c > hi) hi = c; }
        w->mlo[r] = lo; w->mhi[r] = hi;
    }
    /* band limits of (mu/h M - J): only used to skip exact zeros in the LU loops */
    if (pc->problem == PB_NPENDULUM && mass_in == NULL && jac_const == NULL) { w->lu.kl = 9; w->lu.ku = 7; }
    else { w->lu.kl = n; w->lu.ku = n; }

    double t = t0;
    for (int c = 0; c < n; ++c) w->y[c] = y0[(int64_t)c * B + sys];
    prob_rhs(pc, t, w->y, w->f); cnt[C_NFEV] = 1;                    /* :304 */
    double h_abs_state = (o->first_step > 0) ? o->first_step : dir * fmin(1e-6, fabs(tb - t0)); /* :307-315 */
    double h_abs_old_state = 0, err_old_state = 0;
    int hist_state = 0;
    const int fd = has_jac_fn && o->jac_finite_differences;
    if (fd) { for (int c = 0; c < n; ++c) w->jfac[c] = NJ_FACTOR0; num_jac_fd(pc, w, t, atol); cnt[C_NJEV] = 1; } /* :475-481 */
    else if (chain) { chain_jac(pc, w->y, w->ch.nd); cnt[C_NJEV] = 1; }
    else if (has_jac_fn) { prob_jac(pc, t, w->y, w->J); cnt[C_NJEV] = 1; } /* :483-484 */
    else memcpy(w->J, jac_const, sizeof(double) * n * n);
    int current_jac = 1, lu_valid = 0, have_sol = 0;
    double inv_h_sol = 0, t_sol_old = 0, inv_h_lu = 1.0;
    int status = 1; /* running */
    int ei = 0;     /* next t_eval index */
    int64_t nsteps = 0;

    if (t == tb) { /* OdeSolver.step corner case, scipy base.py:193-198 */
        for (; ei < T; ++ei)
            for (int c = 0; c < n; ++c) y_eval[((int64_t)ei * n + c) * B + sys] = w->y[c];
        status = ST_FINISHED;
    }

    while (status == 1) {
        ++nsteps;
        if (o->nmax_step > 0 && nsteps > o->nmax_step) { status = ST_NMAX; break; }
        /* ---------------- prologue, :516-562 */
        cnt[C_NSTEP]++;
        double min_step = fmax(1e-20, 10 * fabs(nextafter(t, dir * INFINITY) - t));
        double h_abs, h_abs_old = h_abs_old_state, err_old = err_old_state;
        int hist = hist_state;
        if (h_abs_state > o->max_step) { h_abs = o->max_step; hist = 0; }
        else if (h_abs_state < min_step) { h_abs = min_step; hist = 0; }
        else h_abs = h_abs_state;
        if (fabs(tb - (t + dir * h_abs)) < 1e-2 * h_abs) { h_abs = fabs(tb - t); lu_valid = 0; }
        int rejected = 0, accepted = 0;
        int n_iter = 0;
        double rate = 0, h = 0, inv_h = 0, t_new = t, error_norm = 0, safety = 0;
        int have_rate = 0;

        while (!accepted) {
            if (h_abs < min_step) { status = ST_TOO_SMALL; break; }
            h = h_abs * dir;
            t_new = t + h;
            if (dir * (t_new - tb) > 0) t_new = tb;
            h = t_new - t;
            h_abs = fabs(h);
            inv_h = 1.0 / h; /* the one division by h of an attempt; mu/h, Z.E/h and x/h use it */
            /* Newton scale, :585-588, kept as a reciprocal */
            for (int c = 0; c < n; ++c) {
                int ve = var_index[c] > 1 ? var_index[c] - 1 : 0;
                double hp = o->scale_newton_norm ? hpow(h, ve) : 1.0;
                w->inv_ns[c] = hp / (atol[c] + fabs(w->y[c]) * rtol);
            }
            int converged = 0, nbad_last = 0;
            for (;;) { /* convergence loop :591-647 */
                /* starting guess :580-583 -> W = TI Z0 (recomputed on a retry: Z0 is never mutated) */
                if (!have_sol || !o->extrapolated_guess) memset(w->W, 0, sizeof(double) * 3 * n);
                else {
                    for (int i = 0; i < 3; ++i) {
                        double tt = t + h * RC[i];
                        double x = (tt - t_sol_old) * inv_h_sol;
                        double p2 = x * x, p3 = p2 * x;
                        for (int c = 0; c < n; ++c) {
                            double v = w->Q[c * 3 + 0] * x;
                            v = fma(w->Q[c * 3 + 1], p2, v);
                            v = fma(w->Q[c * 3 + 2], p3, v);
                            v = v + w->yold[c];
                            w->Z[i * n + c] = v - w->y[c];
                        }
                    }
                    for (int k = 0; k < 3; ++k)
                        for (int c = 0; c < n; ++c) {
                            double v = RTI[k][0] * w->Z[c];
                            v = fma(RTI[k][1], w->Z[n + c], v);
                            w->W[k * n + c] = fma(RTI[k][2], w->Z[2 * n + c], v);
                        }
                }
                if (!lu_valid) { /* :594-613 */
                    if (o->scale_residuals) inv_h_lu = inv_h;
                    for (int c = 0; c < n; ++c)
                        w->hs[c] = (o->scale_residuals && var_index[c] >= 1) ? inv_h_lu : 1.0;
                    double mr = MU_REAL * inv_h, mcr = MU_CR * inv_h, mci = MU_CI * inv_h;
                    if (chain) {
                        if (o->scale_residuals) h_lu = h;
                        mr_f = mr; mcr_f = mcr; mci_f = mci;
                        chain_factor(pc, &w->ch, mr, mcr, mci);
                    } else {
                    for (int r = 0; r < n; ++r)
                        for (int c = 0; c < n; ++c) {
                            double m = w->M[r * n + c], j = w->J[r * n + c], s = w->hs[r];
                            w->lu.lr[r * n + c] = s * fma(mr, m, -j);
                            w->lu.lcr[r * n + c] = s * fma(mcr, m, -j);
                            w->lu.lci[r * n + c] = s * (mci * m);
                        }
                    lu_real(&w->lu);
                    lu_cplx(&w->lu);
                    }
                    cnt[C_NLU]++;
                    trace_ev('F');
                    report(rc, B, sys, RC_FACTORISATION, t, h, -1, -1, -1.0, -1.0);       /* :607 */
                    lu_valid = 1;
                }
                /* ------------- simplified Newton, :793-949 */
                double mr = MU_REAL * inv_h, mcr = MU_CR * inv_h, mci = MU_CI * inv_h;
                double dwn_old = 0;
                int have_old = 0, nbad = 0, k;
                have_rate = 0; rate = 0;
                converged = 0;
                for (k = 0; k < MAXIT; ++k) {
                    int finite = 1;
                    for (int i = 0; i < 3; ++i) {
                        stage_Z(n, w->W, i, w->tmp);
                        for (int c = 0; c < n; ++c) w->tmp[c] = w->y[c] + w->tmp[c];
                        prob_rhs(pc, t + h * RC[i], w->tmp, w->F);
                        for (int c = 0; c < n; ++c) {
                            double Fv = w->F[c];
                            if (!isfinite(Fv)) finite = 0;
                            if (i == 0) {
                                w->fr[c] = RTI[0][0] * Fv; w->fcr[c] = RTI[1][0] * Fv; w->fci[c] = RTI[2][0] * Fv;
                            } else {
                                w->fr[c] = fma(RTI[0][i], Fv, w->fr[c]);
                                w->fcr[c] = fma(RTI[1][i], Fv, w->fcr[c]);
                                w->fci[c] = fma(RTI[2][i], Fv, w->fci[c]);
                            }
                        }
                    }
                    cnt[C_NFEV] += 3;
                    trace_ev('N');
                    if (!finite) break;
                    mass_mv(n, w->M, w->mlo, w->mhi, w->W, w->mw0);
                    mass_mv(n, w->M, w->mlo, w->mhi, w->W + n, w->mw1);
                    mass_mv(n, w->M, w->mlo, w->mhi, w->W + 2 * n, w->mw2);
                    for (int c = 0; c < n; ++c) {
                        double a = fma(-mr, w->mw0[c], w->fr[c]);
                        double br = fma(-mcr, w->mw1[c], w->fcr[c]); br = fma(mci, w->mw2[c], br);
                        double bi = fma(-mcr, w->mw2[c], w->fci[c]); bi = fma(-mci, w->mw1[c], bi);
                        if (chain) { w->fr[c] = a; w->fcr[c] = br; w->fci[c] = bi; }      /* no row scaling */
                        else { w->fr[c] = a * w->hs[c]; w->fcr[c] = br * w->hs[c]; w->fci[c] = bi * w->hs[c]; }
                    }
                    if (chain) {
                        chain_solve_real(pc, &w->ch, mr_f, w->fr);
                        chain_solve_cplx(pc, &w->ch, mcr_f, mci_f, w->fcr, w->fci);
                    } else {
                        solve_real(&w->lu, w->fr);
                        solve_cplx(&w->lu, w->fcr, w->fci);
                    }
                    cnt[C_NSOLVE]++;
                    for (int c = 0; c < n; ++c) {
                        w->sq[c] = w->fr[c] * w->inv_ns[c];
                        w->sq[n + c] = w->fcr[c] * w->inv_ns[c];
                        w->sq[2 * n + c] = w->fci[c] * w->inv_ns[c];
                    }
                    double dwn = sqrt(chain ? chain_sum_squares(w->sq, n / 5, 3) : sum_squares(w->sq, 3 * n, n > 8)) * rsq3n;
                    for (int c = 0; c < n; ++c) {
                        w->W[c] += w->fr[c]; w->W[n + c] += w->fcr[c]; w->W[2 * n + c] += w->fci[c];
                    }
                    double dw_true = 0, dw_true_max = 0;
                    if (have_old) {
                        rate = dwn / dwn_old; have_rate = 1;
                        double inv1 = 1.0 / (1.0 - rate);
                        dw_true = rate * inv1 * dwn;
                        double rp = rate;
                        for (int e = 1; e < MAXIT - k; ++e) rp *= rate;
                        dw_true_max = rp * inv1 * dwn;
                    }
                    if (dwn < tol) { converged = 1; break; }
                    if (have_rate) {
                        if (rate >= 1) {
                            if (rate < 100 && nbad < nmax_bad) { nbad++; continue; }
                            if (!o->all_newton_iterations) break;
                        }
                        if (dw_true < tol) { converged = 1; break; }
                        if (dw_true_max > tol && o->predictive_newton_stop) {
                            if (nbad < nmax_bad) { nbad++; continue; }
                            if (!o->all_newton_iterations) break;
                        }
                    }
                    dwn_old = dwn; have_old = 1;
                }
                n_iter = (k < MAXIT) ? k + 1 : MAXIT;
                nbad_last = nbad;
                safety = o->safety_factor * (2 * MAXIT + 1) / (2 * MAXIT + n_iter); /* :631 */
                if (converged) break;
                if (current_jac) break;                                            /* :638-641 */
                if (fd) num_jac_fd(pc, w, t, atol); else if (chain) chain_jac(pc, w->y, w->ch.nd); else prob_jac(pc, t, w->y, w->J);
                cnt[C_NJEV]++;                                                     /* :643 */
                report(rc, B, sys, RC_JACOBIAN_UPDATE, t, h, -1, -1, -1.0, -1.0);  /* :644 */
                current_jac = 1; lu_valid = 0;
            }
            if (!converged) {                                                      /* :650-658 */
                cnt[C_NFAIL]++;
                report(rc, B, sys, RC_NON_CONV, t, h, n_iter, nbad_last, -1.0, -1.0); /* :654 */
                h_abs *= o->factor_on_non_convergence;
                lu_valid = 0;
                continue;
            }
            for (int i = 0; i < 3; ++i) stage_Z(n, w->W, i, w->Z + i * n);
            trace_ev('E');
            if (o->constant_dt) { accepted = 1; error_norm = 0; break; }
            /* ------------- error estimate, :669-711 */
            for (int c = 0; c < n; ++c) {
                double v = w->Z[c] * RE[0];
                v = fma(w->Z[n + c], RE[1], v);
                v = fma(w->Z[2 * n + c], RE[2], v);
                w->ze[c] = v * inv_h;
                double ynew = w->y[c] + w->Z[2 * n + c];
                int ve = var_index[c] > 1 ? var_index[c] - 1 : 0;
                double hp = o->scale_error ? hpow(h, ve) : 1.0;
                w->inv_es[c] = hp / (atol[c] + fmax(fabs(w->y[c]), fabs(ynew)) * rtol);
            }
            mass_mv(n, w->M, w->mlo, w->mhi, w->ze, w->mw0);
            double err1 = NAN, err2 = NAN;
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 0) for (int c = 0; c < n; ++c) w->err[c] = w->f[c] + w->mw0[c];
                else {
                    for (int c = 0; c < n; ++c) w->tmp[c] = w->y[c] + w->err[c];
                    prob_rhs(pc, t, w->tmp, w->F); cnt[C_NFEV]++;
                    for (int c = 0; c < n; ++c) w->err[c] = w->F[c] + w->mw0[c];
                }
                if (chain) {   /* the reference applies factors of the hscale-d matrix to the unscaled vector */
                    for (int c = 0; c < n; ++c) if (o->scale_residuals && var_index[c] >= 1) w->err[c] = w->err[c] * h_lu;
                    chain_solve_real(pc, &w->ch, mr_f, w->err);
                }
                else solve_real(&w->lu, w->err);    /* rhs deliberately NOT scaled by hscale */
                cnt[C_NSOLVE_ERR]++;
                for (int c = 0; c < n; ++c) {
                    if (o->zero_algebraic_error && var_index[c] != 0) { w->err[c] = 0.0; w->sq[c] = 0.0; continue; }
                    w->sq[c] = w->err[c] * w->inv_es[c];
                }
                error_norm = sqrt(chain ? chain_sum_squares(w->sq, n / 5, 1) : sum_squares(w->sq, n, n > 8)) * (o->zero_algebraic_error ? rsqdiff : rsqn);
                if (pass == 0) err1 = error_norm; else err2 = error_norm;
                if (!(error_norm > 1 && (o->always_2nd_estimate || rejected))) break;
            }
            if (error_norm > 1) {                                                   /* :713-733 */
                double factor = gustafsson(h_abs, h_abs_old, error_norm, err_old, hist, o->predictive_controller);
                report(rc, B, sys, RC_REJECTED, t, h, n_iter, nbad_last, err1, err2);   /* :719 */
                h_abs *= fmax(o->min_factor, safety * factor);
                lu_valid = 0; rejected = 1; cnt[C_NREJ]++;
            } else { report(rc, B, sys, RC_ACCEPTED, t, h, n_iter, nbad_last, err1, err2); cnt[C_NACC]++; accepted = 1; }
        }
        if (status != 1) break;
        /* ---------------- epilogue, :742-784 */
        int recompute = has_jac_fn && n_iter > 2 && have_rate && rate > o->jac_recompute_factor;
        double factor;
        if (o->constant_dt) factor = o->max_step / h_abs;
        else factor = fmin(o->max_factor, safety * gustafsson(h_abs, h_abs_old, error_norm, err_old, hist, o->predictive_controller));
        if (!recompute && factor < 1.2) factor = 1; else lu_valid = 0;
        for (int c = 0; c < n; ++c) { w->yold[c] = w->y[c]; w->y[c] = w->y[c] + w->Z[2 * n + c]; }
        prob_rhs(pc, t_new, w->y, w->f); cnt[C_NFEV]++;
        if (recompute) {
            if (fd) num_jac_fd(pc, w, t_new, atol); else if (chain) chain_jac(pc, w->y, w->ch.nd); else prob_jac(pc, t_new, w->y, w->J);
            cnt[C_NJEV]++; current_jac = 1;
        }
        else if (has_jac_fn) current_jac = 0;
        h_abs_old_state = h_abs_state; err_old_state = error_norm; hist_state = 1;
        h_abs_state = h_abs * factor;
        for (int c = 0; c < n; ++c)
            for (int j = 0; j < 3; ++j) {
                double v = w->Z[c] * RP[0][j];
                v = fma(w->Z[n + c], RP[1][j], v);
                w->Q[c * 3 + j] = fma(w->Z[2 * n + c], RP[2][j], v);
            }
        t_sol_old = t; inv_h_sol = inv_h; have_sol = 1;
        t = t_new;
        /* ---------------- driver: finished? t_eval points inside (t_old, t] */
        if (dir * (t - tb) >= 0) status = ST_FINISHED;
        while (ei < T && (dir > 0 ? t_eval[ei] <= t : t_eval[ei] >= t)) {
            double x = (t_eval[ei] - t_sol_old) * inv_h_sol;
            double p2 = x * x, p3 = p2 * x;
            for (int c = 0; c < n; ++c) {
                double v = w->Q[c * 3 + 0] * x;
                v = fma(w->Q[c * 3 + 1], p2, v);
                v = fma(w->Q[c * 3 + 2], p3, v);
                y_eval[((int64_t)ei * n + c) * B + sys] = v + w->yold[c];
            }
            ++ei;
        }
    }
    for (int e = ei; e < T; ++e)
        for (int c = 0; c < n; ++c) y_eval[((int64_t)e * n + c) * B + sys] = NAN;
    n_eval_out[sys] = ei;
    t_final[sys] = t;
    for (int c = 0; c < n; ++c) y_final[(int64_t)c * B + sys] = w->y[c];
    status_out[sys] = status;
    for (int k = 0; k < C_COUNT; ++k) counters[(int64_t)k * B + sys] = cnt[k];
}

/* Event trace of system `sys` of a batch (see trace_ev); returns the number of events. */
int orc_solve(int problem, int n, int n_params, int64_t B, const double *y0, const double *params,
              const double *mass, const double *jac_const, const int32_t *var_index,
              const double *atol, double rtol, double t0, double t_bound, const double *t_eval,
              int T, const orc_options *opts, double *y_eval, int32_t *n_eval, double *t_final,
              double *y_final, int32_t *status, int32_t *counters, int nthreads);
int orc_trace(int problem, int n, int n_params, const double *y0, const double *params, const int32_t *var_index,
              const double *atol, double rtol, double t0, double t_bound, const orc_options *opts, char *buf, int cap) {
    double *ye = NULL, tf, *yf = malloc(sizeof(double) * n);
    int32_t ne, st, cn[C_COUNT];
    g_trace = buf; g_trace_cap = cap; g_trace_len = 0;
    orc_solve(problem, n, n_params, 1, y0, params, NULL, NULL, var_index, atol, rtol, t0, t_bound, NULL, 0, opts, ye, &ne,
              &tf, yf, &st, cn, 1);
    g_trace = NULL;
    free(yf);
    return g_trace_len;
}

/* Batch entry.  All pointers are host pointers; SoA layouts as in include/brdae.h.
 * mass / jac_const: [n][n] row-major shared by the batch, or NULL (problem's own). */
int orc_solve_report(int problem, int n, int n_params, int64_t B, const double *y0, const double *params,
                     const double *mass, const double *jac_const, const int32_t *var_index,
                     const double *atol, double rtol, double t0, double t_bound, const double *t_eval,
                     int T, const orc_options *opts, double *y_eval, int32_t *n_eval, double *t_final,
                     double *y_final, int32_t *status, int32_t *counters, int nthreads,
                     int report_cap, int32_t *report_count, int32_t *report_code, int32_t *report_it, double *report_val) {
    if (n <= 0 || B < 0) return -1;
    report_ctx rctx = {report_cap, report_count, report_code, report_it, report_val};
    const report_ctx *rc = (report_cap > 0 && report_count && report_code && report_it && report_val) ? &rctx : NULL;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        work w;
        work_alloc(&w, n);
        double *p = xcalloc((size_t)(n_params > 0 ? n_params : 1), 8);
        prob_ctx pc = {problem, n, p, 0, NULL, NULL, NULL};
        if (problem == PB_NPENDULUM) { pc.nn = n / 5; npe_constants(&pc); }
#pragma omp for schedule(dynamic, 16)
        for (int64_t s = 0; s < B; ++s) {
            for (int k = 0; k < n_params; ++k) p[k] = params[(int64_t)k * B + s];
            integrate_one(&pc, &w, opts, mass, jac_const, var_index, atol, rtol, t0, t_bound, t_eval, T, B,
                          s, y0, y_eval, n_eval, t_final, y_final, status, counters, rc);
        }
        free(p); free(pc.m); free(pc.inv_m); free(pc.r0s);
        work_free(&w);
    }
    return 0;
}

int orc_solve(int problem, int n, int n_params, int64_t B, const double *y0, const double *params,
              const double *mass, const double *jac_const, const int32_t *var_index,
              const double *atol, double rtol, double t0, double t_bound, const double *t_eval,
              int T, const orc_options *opts, double *y_eval, int32_t *n_eval, double *t_final,
              double *y_final, int32_t *status, int32_t *counters, int nthreads) {
    return orc_solve_report(problem, n, n_params, B, y0, params, mass, jac_const, var_index, atol, rtol, t0, t_bound, t_eval, T,
                            opts, y_eval, n_eval, t_final, y_final, status, counters, nthreads, 0, NULL, NULL, NULL, NULL);
}
