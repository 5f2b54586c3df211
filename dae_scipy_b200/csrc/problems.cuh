// problems.cuh -- device functors for the reference's test problems (right-hand side, analytic
// Jacobian, mass-matrix policy).  Expression order follows the reference closures:
//   pendulum index 3/2/1 ... scipyDAE/tests/pendulum.py:129-184
//   Robertson DAE .......... scipyDAE/tests/robertson.py:78-87
//   transistor amplifier ... scipyDAE/tests/transistor_amplifier.py:33-56 (Jacobian: analytic
//                            form of the complex-step derivative at :61-65)
// A functor exposes
//   N, NP                          state size, number of per-system parameters
//   struct P; load(P&, params, B, idx)
//   rhs(P, t, y[N], f[N]);  jac(P, t, y[N], J[N][N])
//   using Mass = ...               default mass-matrix policy (see below)
// Mass policies describe a constant matrix M through a compile-time sparsity pattern nz(r,c)
// and a value getter, so that the generic kernel code skips structural zeros at compile time.
#pragma once
#include "radau_common.cuh"

namespace brdae {

// ---------------------------------------------------------------- mass-matrix policies
// diag(b_0..b_{N-1}) with b_r = bit r of ONES (1.0) else 0.0
template <int N_, unsigned ONES>
struct MassDiag01 {
    static constexpr int N = N_;
    static BRDAE_CX bool nz(int r, int c) { return r == c && ((ONES >> r) & 1u); }
    template <class P>
    static BRDAE_HD double get(const P &, const SolveArgs &, int, int) { return 1.0; }
};
// user-supplied dense matrix shared by the ensemble (SolveArgs::mass, row-major)
template <int N_>
struct MassDense {
    static constexpr int N = N_;
    static BRDAE_CX bool nz(int, int) { return true; }
    template <class P>
    static BRDAE_HD double get(const P &, const SolveArgs &a, int r, int c) { return a.mass[r * N + c]; }
};

// M.v : fma chain over the structurally non-zero columns in increasing order (first term a
// plain product) -- the same chain as mass_mv() of the scalar restatement
template <class Mass, class P, int N>
BRDAE_HD void mass_mv(const P &p, const SolveArgs &a, const double (&v)[N], double (&o)[N]) {
#pragma unroll
    for (int r = 0; r < N; ++r) {
        double s = 0.0;
        bool first = true;
#pragma unroll
        for (int c = 0; c < N; ++c) {
            if (Mass::nz(r, c)) {
                double m = Mass::get(p, a, r, c);
                s = first ? m * v[c] : fma(m, v[c], s);
                first = false;
            }
        }
        o[r] = s;
    }
}
template <class Mass, int N>
BRDAE_CX bool mass_row_empty(int r) {
    for (int c = 0; c < N; ++c)
        if (Mass::nz(r, c)) return false;
    return true;
}

// ---------------------------------------------------------------- pendulum (index 3, 2, 1)
template <int INDEX>
struct Pendulum {
    static constexpr bool kStaticLU = true, kFirstIterSkip = true;     // measured: profiles/r01_k1_fastpaths.md
    static constexpr int N = 5, NP = 3;
    // im = 1/m and r0s = r0*r0 are formed once per system (the contract multiplies by 1/m where the
    // reference divides by m; identical for the reference's m = 1)
    struct P { double im, g, r0s; };
    using Mass = MassDiag01<5, 0x0Fu>;  // pendulum.py:139-140
    static BRDAE_HD void load(P &p, const double *params, int64_t B, int64_t i) {
        const double r0 = params[2 * B + i];
        p.im = 1.0 / params[i]; p.g = params[B + i]; p.r0s = r0 * r0;
    }
    static BRDAE_HD void rhs(const P &p, double, const double (&X)[5], double (&f)[5]) {
        const double x = X[0], y = X[1], vx = X[2], vy = X[3], l = X[4];
        f[0] = vx;
        f[1] = vy;
        f[2] = -x * l * p.im;
        f[3] = -p.g - (y * l) * p.im;
        if (INDEX == 3) f[4] = x * x + y * y - p.r0s;
        else if (INDEX == 2) f[4] = x * vx + y * vy;
        else f[4] = l * (x * x + y * y) * p.im + p.g * y - (vx * vx + vy * vy);
    }
    // structural non-zeros of the Jacobian below (radau_k1.cuh: LuPattern)
    static BRDAE_CX bool jac_nz(int r, int c) {
        if (r == 0) return c == 2;
        if (r == 1) return c == 3;
        if (r == 2) return c == 0 || c == 4;
        if (r == 3) return c == 1 || c == 4;
        if (INDEX == 3) return c == 0 || c == 1;
        if (INDEX == 2) return c < 4;
        return true;
    }
    static BRDAE_HD void jac(const P &p, double, const double (&X)[5], double (&J)[5][5]) {
        const double x = X[0], y = X[1], vx = X[2], vy = X[3], l = X[4];
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) J[r][c] = 0.0;
        J[0][2] = 1.0;
        J[1][3] = 1.0;
        J[2][0] = -l * p.im; J[2][4] = -x * p.im;
        J[3][1] = -l * p.im; J[3][4] = -y * p.im;
        if (INDEX == 3) { J[4][0] = 2 * x; J[4][1] = 2 * y; }
        else if (INDEX == 2) { J[4][0] = vx; J[4][1] = vy; J[4][2] = x; J[4][3] = y; }
        else {
            J[4][0] = 2 * x * l * p.im; J[4][1] = 2 * y * l * p.im + p.g; J[4][2] = -2 * vx;
            J[4][3] = -2 * vy; J[4][4] = (x * x + y * y) * p.im;
        }
    }
};

// ---------------------------------------------------------------- Robertson DAE
struct Robertson {
    static constexpr bool kStaticLU = true, kFirstIterSkip = true;
    static constexpr int N = 3, NP = 3;
    struct P { double k1, k2, k3; };
    using Mass = MassDiag01<3, 0x3u>;  // robertson.py:78
    static BRDAE_HD void load(P &p, const double *params, int64_t B, int64_t i) {
        p.k1 = params[i]; p.k2 = params[B + i]; p.k3 = params[2 * B + i];
    }
    static BRDAE_HD void rhs(const P &p, double, const double (&y)[3], double (&f)[3]) {
        f[0] = -p.k1 * y[0] + p.k2 * y[1] * y[2];
        f[1] = p.k1 * y[0] - p.k2 * y[1] * y[2] - p.k3 * (y[1] * y[1]);
        f[2] = y[0] + y[1] + y[2] - 1;
    }
    static BRDAE_HD void jac(const P &p, double, const double (&y)[3], double (&J)[3][3]) {
        J[0][0] = -p.k1; J[0][1] = p.k2 * y[2]; J[0][2] = p.k2 * y[1];
        J[1][0] = p.k1; J[1][1] = -p.k2 * y[2] - 2 * p.k3 * y[1]; J[1][2] = -p.k2 * y[1];
        J[2][0] = 1; J[2][1] = 1; J[2][2] = 1;
    }
};

// ---------------------------------------------------------------- transistor amplifier
struct Transistor {
    static constexpr int N = 5, NP = 12;
    // conductances and the constant source terms are formed once per system (the contract multiplies
    // by 1/R where the reference divides by R)
    struct P { double C1, C2, C3, g0, g12, g3, g4, g5, ub2, ub4, Ua, w; };
    // M = [[C1,-C1,0,0,0],[-C1,C1,0,0,0],[0,0,C2,0,0],[0,0,0,C3,-C3],[0,0,0,-C3,C3]]
    struct Mass {
        static constexpr int N = 5;
        static BRDAE_CX bool nz(int r, int c) {
            return (r < 2 && c < 2) || (r == 2 && c == 2) || (r >= 3 && c >= 3);
        }
        static BRDAE_HD double get(const P &p, const SolveArgs &, int r, int c) {
            if (r < 2) return r == c ? p.C1 : -p.C1;
            if (r == 2) return p.C2;
            return r == c ? p.C3 : -p.C3;
        }
    };
    static BRDAE_HD void load(P &p, const double *q, int64_t B, int64_t i) {
        p.C1 = q[i]; p.C2 = q[B + i]; p.C3 = q[2 * B + i];
        const double R0 = q[3 * B + i], R1 = q[4 * B + i], R2 = q[5 * B + i], R3 = q[6 * B + i], R4 = q[7 * B + i];
        const double R5 = q[8 * B + i], Ub = q[9 * B + i];
        p.g0 = 1 / R0; p.g12 = 1 / R1 + 1 / R2; p.g3 = 1 / R3; p.g4 = 1 / R4; p.g5 = 1 / R5;
        p.ub2 = Ub / R2; p.ub4 = Ub / R4;
        p.Ua = q[10 * B + i]; p.w = q[11 * B + i];
    }
    static BRDAE_HD void rhs(const P &p, double t, const double (&y)[5], double (&f)[5]) {
        const double Ue = p.Ua * sin(p.w * t);
        const double fT = 1e-6 * (exp((y[1] - y[2]) * (1 / 0.026)) - 1);
        f[0] = (Ue - y[0]) * p.g0;
        f[1] = p.ub2 - y[1] * p.g12 - 0.01 * fT;
        f[2] = fT - y[2] * p.g3;
        f[3] = p.ub4 - y[3] * p.g4 - 0.99 * fT;
        f[4] = -y[4] * p.g5;
    }
    static BRDAE_HD void jac(const P &p, double, const double (&y)[5], double (&J)[5][5]) {
        const double df = 1e-6 * exp((y[1] - y[2]) * (1 / 0.026)) * (1 / 0.026);
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) J[r][c] = 0.0;
        J[0][0] = -p.g0;
        J[1][1] = -p.g12 - 0.01 * df; J[1][2] = 0.01 * df;
        J[2][1] = df; J[2][2] = -df - p.g3;
        J[3][1] = -0.99 * df; J[3][2] = 0.99 * df; J[3][3] = -p.g4;
        J[4][4] = -p.g5;
    }
};

// ---------------------------------------------------------------- Robertson, ODE formulation
// robertson.py:69-77; used by the reference's DAE-vs-ODE cross-check (:249-251)
struct RobertsonODE {
    static constexpr bool kStaticLU = true, kFirstIterSkip = true;
    static constexpr int N = 3, NP = 3;
    struct P { double k1, k2, k3; };
    using Mass = MassDiag01<3, 0x7u>;
    static BRDAE_HD void load(P &p, const double *params, int64_t B, int64_t i) {
        p.k1 = params[i]; p.k2 = params[B + i]; p.k3 = params[2 * B + i];
    }
    static BRDAE_HD void rhs(const P &p, double, const double (&y)[3], double (&f)[3]) {
        f[0] = -p.k1 * y[0] + p.k2 * y[1] * y[2];
        f[1] = p.k1 * y[0] - p.k2 * y[1] * y[2] - p.k3 * (y[1] * y[1]);
        f[2] = p.k3 * (y[1] * y[1]);
    }
    static BRDAE_HD void jac(const P &p, double, const double (&y)[3], double (&J)[3][3]) {
        J[0][0] = -p.k1; J[0][1] = p.k2 * y[2]; J[0][2] = p.k2 * y[1];
        J[1][0] = p.k1; J[1][1] = -p.k2 * y[2] - 2 * p.k3 * y[1]; J[1][2] = -p.k2 * y[1];
        J[2][0] = 0; J[2][1] = 2 * p.k3 * y[1]; J[2][2] = 0;
    }
};

// ---------------------------------------------------------------- Jay 1995 index-3 DAE
// jay_dae.py:22-68; parameter: nonlinear-multiplier flag.  Analytic Jacobian (FD in the reference).
struct Jay {
    static constexpr int N = 5, NP = 1;
    struct P { double nonlinear; };
    using Mass = MassDiag01<5, 0x0Fu>;
    static BRDAE_HD void load(P &p, const double *params, int64_t, int64_t i) { p.nonlinear = params[i]; }
    static BRDAE_HD void rhs(const P &p, double, const double (&y)[5], double (&f)[5]) {
        const double y1 = y[0], y2 = y[1], z1 = y[2], z2 = y[3], la = y[4];
        f[0] = 2 * y1 * y2 * z1 * z2;
        f[1] = -y1 * y2 * (z2 * z2);
        f[2] = (y1 * y2 + z1 * z2) * la;
        if (p.nonlinear != 0.0) f[3] = -y1 * (y2 * y2) * (z2 * z2 * z2) * (la * la);
        else f[3] = -y1 * (y2 * y2) * (z2 * z2) * la;
        f[4] = y1 * (y2 * y2) - 1.0;
    }
    static BRDAE_HD void jac(const P &p, double, const double (&y)[5], double (&J)[5][5]) {
        const double y1 = y[0], y2 = y[1], z1 = y[2], z2 = y[3], la = y[4];
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) J[r][c] = 0.0;
        J[0][0] = 2 * y2 * z1 * z2; J[0][1] = 2 * y1 * z1 * z2; J[0][2] = 2 * y1 * y2 * z2; J[0][3] = 2 * y1 * y2 * z1;
        J[1][0] = -y2 * (z2 * z2); J[1][1] = -y1 * (z2 * z2); J[1][3] = -2 * y1 * y2 * z2;
        J[2][0] = y2 * la; J[2][1] = y1 * la; J[2][2] = z2 * la; J[2][3] = z1 * la; J[2][4] = y1 * y2 + z1 * z2;
        if (p.nonlinear != 0.0) {
            J[3][0] = -(y2 * y2) * (z2 * z2 * z2) * (la * la); J[3][1] = -2 * y1 * y2 * (z2 * z2 * z2) * (la * la);
            J[3][3] = -3 * y1 * (y2 * y2) * (z2 * z2) * (la * la); J[3][4] = -2 * y1 * (y2 * y2) * (z2 * z2 * z2) * la;
        } else {
            J[3][0] = -(y2 * y2) * (z2 * z2) * la; J[3][1] = -2 * y1 * y2 * (z2 * z2) * la;
            J[3][3] = -2 * y1 * (y2 * y2) * z2 * la; J[3][4] = -y1 * (y2 * y2) * (z2 * z2);
        }
        J[4][0] = y2 * y2; J[4][1] = 2 * y1 * y2;
    }
};

// ---------------------------------------------------------------- polynomial index-2 DAE
// validation_dense_output.py:23-59 with p = 2: y' = P'(t) + (z - y), 0 = y - P(t);
// parameters (a4, a3, a2, a1, a0), Horner evaluation
struct PolyDAE2 {
    static constexpr int N = 2, NP = 5;
    struct P { double a[5]; };
    using Mass = MassDiag01<2, 0x1u>;
    static BRDAE_HD void load(P &p, const double *params, int64_t B, int64_t i) {
#pragma unroll
        for (int k = 0; k < 5; ++k) p.a[k] = params[k * B + i];
    }
    static BRDAE_HD void rhs(const P &p, double t, const double (&x)[2], double (&f)[2]) {
        const double Pv = (((p.a[0] * t + p.a[1]) * t + p.a[2]) * t + p.a[3]) * t + p.a[4];
        const double dP = ((4 * p.a[0] * t + 3 * p.a[1]) * t + 2 * p.a[2]) * t + p.a[3];
        f[0] = dP + (x[1] - x[0]);
        f[1] = x[0] - Pv;
    }
    static BRDAE_HD void jac(const P &, double, const double (&)[2], double (&J)[2][2]) {
        J[0][0] = -1; J[0][1] = 1; J[1][0] = 1; J[1][1] = 0;
    }
};

// ---------------------------------------------------------------- linear ODE (mass-matrix check)
// linear_ODE_with_mass_matrix.py:32-39: f = c * x componentwise, 7 unknowns, parameters c[7]
struct Linear7 {
    static constexpr int N = 7, NP = 7;
    struct P { double c[7]; };
    using Mass = MassDiag01<7, 0x7Fu>;
    static BRDAE_HD void load(P &p, const double *params, int64_t B, int64_t i) {
#pragma unroll
        for (int k = 0; k < 7; ++k) p.c[k] = params[k * B + i];
    }
    static BRDAE_HD void rhs(const P &p, double, const double (&x)[7], double (&f)[7]) {
#pragma unroll
        for (int k = 0; k < 7; ++k) f[k] = p.c[k] * x[k];
    }
    static BRDAE_HD void jac(const P &p, double, const double (&)[7], double (&J)[7][7]) {
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int c = 0; c < 7; ++c) J[r][c] = (r == c) ? p.c[r] : 0.0;
    }
};

}  // namespace brdae
