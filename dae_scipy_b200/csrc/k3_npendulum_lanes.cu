// k3_npendulum_lanes.cu -- K3: lane-per-rope batched RadauDAE for LARGE ensembles of n-pendulum
// chains (reference problem scipyDAE/tests/recursive_pendulum.py:93-127), sm_100a.
//
// K2 (one CTA per rope) keeps a rope's state in shared memory, but the banded LU and the
// triangular solves are sequential along the band, so most of its lanes wait (profiles/:
// 57 % barrier stalls, 1 % of FP64 peak).  With thousands of ropes there is a better source of
// parallelism: the ropes themselves.  Here ONE LANE integrates one rope, exactly like the
// thread-per-system kernel K1, except that a rope's state (17 N doubles of vectors and 78 N doubles
// of banded LU factors) lives in global memory, slot-interleaved -- element e of the lane with
// global id g is at [e * G + g] -- so that the 32 lanes of a warp, which walk the band in lockstep,
// always touch 32 consecutive doubles: every access is a fully coalesced 256-byte line.  The
// working set that is hot at any time (a 10 x 26 window of the band per lane) stays in L2; the rest
// streams through HBM.  The bound is L2/HBM bandwidth and latency, not the FP64 pipe.
//
// Control flow is K1's: per-lane state machine (SETUP / FACTOR / NEWTON / ERREST), one warp-wide
// vote per trip, only the block most lanes wait for is executed.  Arithmetic is the contract of
// K2 and of the scalar restatement (banded right-looking LU, first-maximum partial pivoting,
// reciprocal pivots, column-oriented substitutions, norms as 32 interleaved chains + tree), so all
// three agree bit for bit.
//
// Reference map: radauDAE.py:515-791 (step), 793-949 (Newton), 951-982 (dense output),
// 1079-1157 (driver / t_eval).
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace brdae {

namespace {

constexpr int KL = 9, KU = 7, WB = 2 * KL + KU + 1;
constexpr int K3_BLOCK = 32;
constexpr double GRAV = 9.81;

enum : int { S_IDLE = 0, S_SETUP, S_FACTOR, S_NEWTON, S_ERREST, S_EXIT };
enum : int { SETUP_STEP = 0, SETUP_ATTEMPT = 1, SETUP_GUESS = 2 };

struct K3Cfg {
    double *state;      // [95 N][G] doubles
    int *ipiv;          // [2 N][G]
    long long G;        // number of lane slots
};

// slot-interleaved view of one lane's array
struct Vec {
    double *p;
    long long G;
    __device__ __forceinline__ double &operator[](int e) const { return p[(long long)e * G]; }
};
struct IVec {
    int *p;
    long long G;
    __device__ __forceinline__ int &operator[](int e) const { return p[(long long)e * G]; }
};
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
#define ABI(r, c) ((r) * WB + ((c) - (r) + KL))

struct Consts { int N, nn; double m_in, im_in, r0s; };

// f = rhs(Y), nodes walked from the last to the first (recursive_pendulum.py:93-118).
// Yf(c) returns Y[c]; Ff(c, v) consumes f[c].
template <class YF, class FF>
__device__ __forceinline__ void rope_rhs(const Consts &k, YF &&Yf, FF &&Ff) {
    double a_prev = 0, b_prev = 0;
    double xi = Yf(5 * (k.nn - 1)), yi = Yf(5 * (k.nn - 1) + 1);
    for (int i = k.nn - 1; i >= 0; --i) {
        const double vx = Yf(5 * i + 2), vy = Yf(5 * i + 3), la = Yf(5 * i + 4);
        double xm = 0, ym = 0;
        if (i > 0) { xm = Yf(5 * i - 5); ym = Yf(5 * i - 4); }
        const double dx = (i == 0) ? xi : xi - xm;
        const double dy = (i == 0) ? yi : yi - ym;
        const double rs = dx * dx + dy * dy;
        const double aa = la * dx / rs, bb = la * dy / rs;
        const bool last = (i == k.nn - 1);
        const double m = last ? 1.0 : k.m_in, im = last ? 1.0 : k.im_in;
        Ff(5 * i, vx);
        Ff(5 * i + 1, vy);
        if (last) { Ff(5 * i + 2, im * aa); Ff(5 * i + 3, im * (bb - m * GRAV)); }
        else { Ff(5 * i + 2, im * (aa - a_prev)); Ff(5 * i + 3, im * (bb - b_prev - m * GRAV)); }
        Ff(5 * i + 4, rs - k.r0s);
        a_prev = aa; b_prev = bb; xi = xm; yi = ym;
    }
}

struct NodeJac { double adx, ady, bdx, bdy, al, bl, dx, dy; };
__device__ __forceinline__ NodeJac node_jac(const Vec &X, int i) {
    NodeJac j;
    const double x = X[5 * i], y = X[5 * i + 1], l = X[5 * i + 4];
    j.dx = (i == 0) ? x : x - X[5 * i - 5];
    j.dy = (i == 0) ? y : y - X[5 * i - 4];
    const double rs = j.dx * j.dx + j.dy * j.dy, rs2 = rs * rs;
    j.adx = l * (rs - 2 * j.dx * j.dx) / rs2;
    j.ady = l * (-2 * j.dx * j.dy) / rs2;
    j.bdy = l * (rs - 2 * j.dy * j.dy) / rs2;
    j.bdx = j.ady;
    j.al = j.dx / rs;
    j.bl = j.dy / rs;
    return j;
}

// ---- banded LU and solves, one lane each, factors in slot-interleaved global memory -----------
// A lane is alone with its memory latency (there are only B/32 warps), so every loop is written as
// "issue a batch of independent loads into registers, compute, store": the pivot column, the pivot
// row and three target rows at a time in the factorisation; a sliding register window of the
// right-hand side plus a one-column prefetch of the factor in the substitutions.
constexpr int UW = KU + KL;          // 16 columns right of the pivot

__device__ void lu_band_real(const Vec &ab, const IVec &piv, int N) {
    for (int k = 0; k < N; ++k) {
        const int rmax = imin(N - 1, k + KL), cmax = imin(N - 1, k + UW);
        double cv[KL + 1];                         // pivot column, rows k .. k+9
#pragma unroll
        for (int i = 0; i <= KL; ++i) cv[i] = (k + i <= rmax) ? ab[ABI(k + i, k)] : 0.0;
        int p = k;
        double best = fabs(cv[0]);
#pragma unroll
        for (int i = 1; i <= KL; ++i) {
            const double v = fabs(cv[i]);
            if (k + i <= rmax && v > best) { best = v; p = k + i; }
        }
        piv[k] = p;
        double prow[UW + 1], krow[UW + 1];         // new pivot row (old row p) and old row k, columns k .. k+16
#pragma unroll
        for (int j = 0; j <= UW; ++j) {
            const bool in = (k + j <= cmax);
            prow[j] = in ? ab[ABI(p, k + j)] : 0.0;
            krow[j] = in ? ab[ABI(k, k + j)] : 0.0;
        }
        const double rcp = 1.0 / prow[0];
        if (p != k) {
#pragma unroll
            for (int j = 0; j <= UW; ++j) if (k + j <= cmax) { ab[ABI(p, k + j)] = krow[j]; ab[ABI(k, k + j)] = prow[j]; }
        }
        ab[ABI(k, k)] = rcp;
        double l[KL + 1];                          // multipliers of rows k+1 .. k+9 (row p now holds old row k)
#pragma unroll
        for (int i = 1; i <= KL; ++i) {
            l[i] = ((k + i == p) ? cv[0] : cv[i]) * rcp;
            if (k + i <= rmax) ab[ABI(k + i, k)] = l[i];
        }
#pragma unroll
        for (int i0 = 1; i0 <= KL; i0 += 3) {      // three target rows per batch
            double rw[3][UW];
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int j = 1; j <= UW; ++j)
                    rw[q][j - 1] = (k + i0 + q <= rmax && k + j <= cmax) ? ab[ABI(k + i0 + q, k + j)] : 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int j = 1; j <= UW; ++j)
                    if (k + i0 + q <= rmax && k + j <= cmax) ab[ABI(k + i0 + q, k + j)] = fma(-l[i0 + q], prow[j], rw[q][j - 1]);
        }
    }
}
__device__ void lu_band_cplx(const Vec &ar, const Vec &ai, const IVec &piv, int N) {
    for (int k = 0; k < N; ++k) {
        const int rmax = imin(N - 1, k + KL), cmax = imin(N - 1, k + UW);
        double cvr[KL + 1], cvi[KL + 1];
#pragma unroll
        for (int i = 0; i <= KL; ++i) {
            const bool in = (k + i <= rmax);
            cvr[i] = in ? ar[ABI(k + i, k)] : 0.0;
            cvi[i] = in ? ai[ABI(k + i, k)] : 0.0;
        }
        int p = k;
        double best = fabs(cvr[0]) + fabs(cvi[0]);
#pragma unroll
        for (int i = 1; i <= KL; ++i) {
            const double v = fabs(cvr[i]) + fabs(cvi[i]);
            if (k + i <= rmax && v > best) { best = v; p = k + i; }
        }
        piv[k] = p;
        double pr_[UW + 1], pi_[UW + 1], kr_[UW + 1], ki_[UW + 1];
#pragma unroll
        for (int j = 0; j <= UW; ++j) {
            const bool in = (k + j <= cmax);
            pr_[j] = in ? ar[ABI(p, k + j)] : 0.0; pi_[j] = in ? ai[ABI(p, k + j)] : 0.0;
            kr_[j] = in ? ar[ABI(k, k + j)] : 0.0; ki_[j] = in ? ai[ABI(k, k + j)] : 0.0;
        }
        const double invd = 1.0 / fma(pr_[0], pr_[0], pi_[0] * pi_[0]);
        const double qr = pr_[0] * invd, qi = -(pi_[0] * invd);
        if (p != k) {
#pragma unroll
            for (int j = 0; j <= UW; ++j)
                if (k + j <= cmax) {
                    ar[ABI(p, k + j)] = kr_[j]; ai[ABI(p, k + j)] = ki_[j];
                    ar[ABI(k, k + j)] = pr_[j]; ai[ABI(k, k + j)] = pi_[j];
                }
        }
        ar[ABI(k, k)] = qr; ai[ABI(k, k)] = qi;
        double lr[KL + 1], li[KL + 1];
#pragma unroll
        for (int i = 1; i <= KL; ++i) {
            const double xr0 = (k + i == p) ? cvr[0] : cvr[i], xi0 = (k + i == p) ? cvi[0] : cvi[i];
            lr[i] = fma(xr0, qr, -(xi0 * qi));
            li[i] = fma(xr0, qi, xi0 * qr);
            if (k + i <= rmax) { ar[ABI(k + i, k)] = lr[i]; ai[ABI(k + i, k)] = li[i]; }
        }
#pragma unroll
        for (int i0 = 1; i0 <= KL; i0 += 3) {
            double wr[3][UW], wi[3][UW];
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int j = 1; j <= UW; ++j) {
                    const bool in = (k + i0 + q <= rmax && k + j <= cmax);
                    wr[q][j - 1] = in ? ar[ABI(k + i0 + q, k + j)] : 0.0;
                    wi[q][j - 1] = in ? ai[ABI(k + i0 + q, k + j)] : 0.0;
                }
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int j = 1; j <= UW; ++j)
                    if (k + i0 + q <= rmax && k + j <= cmax) {
                        double xr = wr[q][j - 1], xi = wi[q][j - 1];
                        xr = fma(-lr[i0 + q], pr_[j], xr); xr = fma(li[i0 + q], pi_[j], xr);
                        xi = fma(-lr[i0 + q], pi_[j], xi); xi = fma(-li[i0 + q], pr_[j], xi);
                        ar[ABI(k + i0 + q, k + j)] = xr; ai[ABI(k + i0 + q, k + j)] = xi;
                    }
        }
    }
}
__device__ void solve_band_real(const Vec &ab, const IVec &piv, const Vec &b, int N) {
    // forward: window w[i] = b[k+i], i = 0..9; the multipliers of the next column are prefetched
    {
        double w[KL + 1], Lc[KL], Ln[KL];
#pragma unroll
        for (int i = 0; i <= KL; ++i) w[i] = (i < N) ? b[i] : 0.0;
#pragma unroll
        for (int i = 1; i <= KL; ++i) Lc[i - 1] = (i < N) ? ab[ABI(i, 0)] : 0.0;
        int p = piv[0];
        for (int k = 0; k < N; ++k) {
            const int pn = (k + 1 < N) ? piv[k + 1] : 0;
#pragma unroll
            for (int i = 1; i <= KL; ++i) Ln[i - 1] = (k + 1 + i < N) ? ab[ABI(k + 1 + i, k + 1)] : 0.0;
            const double bnew = (k + KL + 1 < N) ? b[k + KL + 1] : 0.0;
            const int d = p - k;
#pragma unroll
            for (int i = 1; i <= KL; ++i) { const bool sw = (d == i); const double u = w[0], v = w[i]; w[0] = sw ? v : u; w[i] = sw ? u : v; }
            const double bk = w[0];
            b[k] = bk;
#pragma unroll
            for (int i = 1; i <= KL; ++i) w[i] = fma(-Lc[i - 1], bk, w[i]);
#pragma unroll
            for (int i = 0; i < KL; ++i) { w[i] = w[i + 1]; Lc[i] = Ln[i]; }
            w[KL] = bnew;
            p = pn;
        }
    }
    // backward, column oriented: window w[j] = b[k-j], j = 0..16
    {
        double w[UW + 1], Uc[UW], Un[UW];
#pragma unroll
        for (int j = 0; j <= UW; ++j) w[j] = (N - 1 - j >= 0) ? b[N - 1 - j] : 0.0;
#pragma unroll
        for (int j = 1; j <= UW; ++j) Uc[j - 1] = (N - 1 - j >= 0) ? ab[ABI(N - 1 - j, N - 1)] : 0.0;
        double dg = ab[ABI(N - 1, N - 1)];
        for (int k = N - 1; k >= 0; --k) {
#pragma unroll
            for (int j = 1; j <= UW; ++j) Un[j - 1] = (k - 1 - j >= 0) ? ab[ABI(k - 1 - j, k - 1)] : 0.0;
            const double dgn = (k >= 1) ? ab[ABI(k - 1, k - 1)] : 0.0;
            const double bnew = (k - UW - 1 >= 0) ? b[k - UW - 1] : 0.0;
            const double xk = w[0] * dg;
            b[k] = xk;
#pragma unroll
            for (int j = 1; j <= UW; ++j) w[j] = fma(-Uc[j - 1], xk, w[j]);
#pragma unroll
            for (int j = 0; j < UW; ++j) { w[j] = w[j + 1]; Uc[j] = Un[j]; }
            w[UW] = bnew;
            dg = dgn;
        }
    }
}
__device__ void solve_band_cplx(const Vec &ar, const Vec &ai, const IVec &piv, const Vec &br, const Vec &bi, int N) {
    {
        double wr[KL + 1], wi[KL + 1], Lr[KL], Li[KL], Lrn[KL], Lin[KL];
#pragma unroll
        for (int i = 0; i <= KL; ++i) { wr[i] = (i < N) ? br[i] : 0.0; wi[i] = (i < N) ? bi[i] : 0.0; }
#pragma unroll
        for (int i = 1; i <= KL; ++i) { Lr[i - 1] = (i < N) ? ar[ABI(i, 0)] : 0.0; Li[i - 1] = (i < N) ? ai[ABI(i, 0)] : 0.0; }
        int p = piv[0];
        for (int k = 0; k < N; ++k) {
            const int pn = (k + 1 < N) ? piv[k + 1] : 0;
#pragma unroll
            for (int i = 1; i <= KL; ++i) {
                const bool in = (k + 1 + i < N);
                Lrn[i - 1] = in ? ar[ABI(k + 1 + i, k + 1)] : 0.0;
                Lin[i - 1] = in ? ai[ABI(k + 1 + i, k + 1)] : 0.0;
            }
            const bool more = (k + KL + 1 < N);
            const double bnr = more ? br[k + KL + 1] : 0.0, bni = more ? bi[k + KL + 1] : 0.0;
            const int d = p - k;
#pragma unroll
            for (int i = 1; i <= KL; ++i) {
                const bool sw = (d == i);
                const double u = wr[0], v = wr[i], ui = wi[0], vi = wi[i];
                wr[0] = sw ? v : u; wr[i] = sw ? u : v; wi[0] = sw ? vi : ui; wi[i] = sw ? ui : vi;
            }
            const double xr = wr[0], xi = wi[0];
            br[k] = xr; bi[k] = xi;
#pragma unroll
            for (int i = 1; i <= KL; ++i) {
                double yr = wr[i], yi = wi[i];
                yr = fma(-Lr[i - 1], xr, yr); yr = fma(Li[i - 1], xi, yr);
                yi = fma(-Lr[i - 1], xi, yi); yi = fma(-Li[i - 1], xr, yi);
                wr[i] = yr; wi[i] = yi;
            }
#pragma unroll
            for (int i = 0; i < KL; ++i) { wr[i] = wr[i + 1]; wi[i] = wi[i + 1]; Lr[i] = Lrn[i]; Li[i] = Lin[i]; }
            wr[KL] = bnr; wi[KL] = bni;
            p = pn;
        }
    }
    {
        double wr[UW + 1], wi[UW + 1], Ur[UW], Ui[UW], Urn[UW], Uin[UW];
#pragma unroll
        for (int j = 0; j <= UW; ++j) { const bool in = (N - 1 - j >= 0); wr[j] = in ? br[N - 1 - j] : 0.0; wi[j] = in ? bi[N - 1 - j] : 0.0; }
#pragma unroll
        for (int j = 1; j <= UW; ++j) {
            const bool in = (N - 1 - j >= 0);
            Ur[j - 1] = in ? ar[ABI(N - 1 - j, N - 1)] : 0.0; Ui[j - 1] = in ? ai[ABI(N - 1 - j, N - 1)] : 0.0;
        }
        double qr = ar[ABI(N - 1, N - 1)], qi = ai[ABI(N - 1, N - 1)];
        for (int k = N - 1; k >= 0; --k) {
#pragma unroll
            for (int j = 1; j <= UW; ++j) {
                const bool in = (k - 1 - j >= 0);
                Urn[j - 1] = in ? ar[ABI(k - 1 - j, k - 1)] : 0.0; Uin[j - 1] = in ? ai[ABI(k - 1 - j, k - 1)] : 0.0;
            }
            const double qrn = (k >= 1) ? ar[ABI(k - 1, k - 1)] : 0.0, qin = (k >= 1) ? ai[ABI(k - 1, k - 1)] : 0.0;
            const bool more = (k - UW - 1 >= 0);
            const double bnr = more ? br[k - UW - 1] : 0.0, bni = more ? bi[k - UW - 1] : 0.0;
            const double sr = wr[0], si = wi[0];
            const double xr = fma(sr, qr, -(si * qi));
            const double xi = fma(sr, qi, si * qr);
            br[k] = xr; bi[k] = xi;
#pragma unroll
            for (int j = 1; j <= UW; ++j) {
                double yr = wr[j], yi = wi[j];
                yr = fma(-Ur[j - 1], xr, yr); yr = fma(Ui[j - 1], xi, yr);
                yi = fma(-Ur[j - 1], xi, yi); yi = fma(-Ui[j - 1], xr, yi);
                wr[j] = yr; wi[j] = yi;
            }
#pragma unroll
            for (int j = 0; j < UW; ++j) { wr[j] = wr[j + 1]; wi[j] = wi[j + 1]; Ur[j] = Urn[j]; Ui[j] = Uin[j]; }
            wr[UW] = bnr; wi[UW] = bni;
            qr = qrn; qi = qin;
        }
    }
}

// sum of squares of v[e] * s[e mod N], e < m: 32 interleaved chains folded by the tree 16,8,4,2,1
__device__ double sumsq_scaled(const Vec &v, const Vec &s, int N, int m) {
    double p[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) p[j] = 0.0;
    for (int e0 = 0; e0 < m; e0 += 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int e = e0 + j;
            if (e < m) { const double x = v[e] * s[e % N]; p[j] = fma(x, x, p[j]); }
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
        for (int j = 0; j < off; ++j) p[j] = p[j] + p[j + off];
    return p[0];
}

__device__ __forceinline__ long long fetch_rope(bool want, unsigned long long *counter) {
    const unsigned need = __ballot_sync(0xffffffffu, want);
    long long idx = -1;
    if (need) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(need) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(need));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (want) idx = (long long)(base + __popc(need & ((1u << lane) - 1u)));
    }
    return idx;
}

__global__ void __launch_bounds__(K3_BLOCK) k3_kernel(const __grid_constant__ SolveArgs a, const K3Cfg cfg) {
    const int N = a.n, nn = N / 5;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x, G = cfg.G;
    Consts kc;
    kc.N = N; kc.nn = nn; kc.m_in = 1.0 / (nn - 1); kc.im_in = 1 / kc.m_in;
    { const double r0 = 2.0 / nn; kc.r0s = r0 * r0; }
    // slot-interleaved arrays of this lane
    double *base = cfg.state + g;
    auto vec = [&](long long first) { return Vec{base + first * G, G}; };
    const Vec y = vec(0), f = vec(N), yold = vec(2LL * N), yJ = vec(3LL * N), Q = vec(4LL * N), W = vec(7LL * N),
              R = vec(10LL * N), inv_ns = vec(13LL * N), inv_es = vec(14LL * N), ze = vec(15LL * N), err = vec(16LL * N),
              lur = vec(17LL * N), lcr = vec(17LL * N + (long long)N * WB), lci = vec(17LL * N + 2LL * N * WB);
    const IVec pivr{cfg.ipiv + g, G}, pivc{cfg.ipiv + (long long)N * G + g, G};
    const Vec R1{R.p + (long long)N * G, G}, R2{R.p + 2LL * N * G, G};
    const double *atol = a.atol_dev;
    const int *vidx = a.var_index_dev;

    const int MAXIT = a.maxit;
    const double dir = a.dir;

    int state = S_IDLE, setup_kind = SETUP_STEP;
    long long idx = -1;
    double t = 0, h_abs_state = 0, h_abs_old_state = 0, err_old_state = 0;
    bool hist_state = false, current_jac = true, lu_valid = false, have_sol = false;
    double inv_h_sol = 1, t_sol_old = 0, inv_h_lu = 1.0;
    int cnt[CN_COUNT];
    int ei = 0;
    long long nsteps = 0;
    double min_step = 0, h_abs = 0, h_abs_old = 0, err_old = 0;
    bool hist = false, rejected = false;
    double h = 0, inv_h = 0, t_new = 0, mr = 0, mcr = 0, mci = 0;
    int k = 0, nbad = 0, n_iter = 0;
    double dwn_old = 0, rate = 0, safety = 0;
    bool have_old = false, have_rate = false;
#pragma unroll
    for (int q = 0; q < CN_COUNT; ++q) cnt[q] = 0;

    for (;;) {
        int finish_status = 1;
        // ---------------- refill
        if (__any_sync(0xffffffffu, state == S_IDLE)) {
            const long long got = fetch_rope(state == S_IDLE, a.work_counter);
            if (state == S_IDLE) {
                if (got >= a.B || g >= G) state = S_EXIT;
                else {
                    idx = got;
                    t = a.t0;
                    for (int c = 0; c < N; ++c) { const double v = a.y0[(long long)c * a.B + idx]; y[c] = v; yJ[c] = v; }
                    rope_rhs(kc, [&](int c) { return y[c]; }, [&](int c, double v) { f[c] = v; });
#pragma unroll
                    for (int q = 0; q < CN_COUNT; ++q) cnt[q] = 0;
                    cnt[CN_NFEV] = 1; cnt[CN_NJEV] = 1;
                    h_abs_state = a.h0;
                    hist_state = false; current_jac = true; lu_valid = false; have_sol = false;
                    inv_h_lu = 1.0; ei = 0; nsteps = 0;
                    state = S_SETUP; setup_kind = SETUP_STEP;
                    if (t == a.tb) {
                        for (; ei < a.T; ++ei)
                            for (int c = 0; c < N; ++c) a.y_eval[((long long)ei * N + c) * a.B + idx] = y[c];
                        finish_status = BRDAE_STATUS_FINISHED;
                        state = S_IDLE;
                    }
                }
            }
            if (__all_sync(0xffffffffu, state == S_EXIT)) break;
        }
        __syncwarp();
        const int nF = __popc(__ballot_sync(0xffffffffu, state == S_FACTOR));
        const int nN = __popc(__ballot_sync(0xffffffffu, state == S_NEWTON));
        const int nE = __popc(__ballot_sync(0xffffffffu, state == S_ERREST));
        const int run = (nN > 0 && nN >= nE && nN >= nF) ? S_NEWTON : ((nE > 0 && nE >= nF) ? S_ERREST : (nF > 0 ? S_FACTOR : S_IDLE));

        if (run == S_FACTOR) {
            if (state == S_FACTOR) {          // iteration matrices + both LUs (:594-613)
                if (a.scale_residuals) inv_h_lu = inv_h;
                for (int e = 0; e < N * WB; ++e) { lur[e] = 0.0; lcr[e] = 0.0; lci[e] = 0.0; }
                for (int i = 0; i < nn; ++i) {
                    const NodeJac me = node_jac(yJ, i);
                    const int r = 5 * i;
                    const double im = (i == nn - 1) ? 1.0 : kc.im_in;
                    auto put = [&](int row, int col, double m, double jv) {
                        const double hsc = (a.scale_residuals && vidx[row] >= 1) ? inv_h_lu : 1.0;
                        lur[ABI(row, col)] = hsc * fma(mr, m, -jv);
                        lcr[ABI(row, col)] = hsc * fma(mcr, m, -jv);
                        lci[ABI(row, col)] = hsc * (mci * m);
                    };
                    put(r, r, 1.0, 0.0); put(r + 1, r + 1, 1.0, 0.0); put(r + 2, r + 2, 1.0, 0.0); put(r + 3, r + 3, 1.0, 0.0);
                    put(r, r + 2, 0.0, 1.0);
                    put(r + 1, r + 3, 0.0, 1.0);
                    double j20 = im * me.adx, j21 = im * me.ady, j30 = im * me.bdx, j31 = im * me.bdy;
                    if (i < nn - 1) {
                        const NodeJac nx = node_jac(yJ, i + 1);
                        j20 = j20 + im * nx.adx; j21 = j21 + im * nx.ady; j30 = j30 + im * nx.bdx; j31 = j31 + im * nx.bdy;
                        put(r + 2, r + 5, 0.0, -(im * nx.adx)); put(r + 2, r + 6, 0.0, -(im * nx.ady)); put(r + 2, r + 9, 0.0, -(im * nx.al));
                        put(r + 3, r + 5, 0.0, -(im * nx.bdx)); put(r + 3, r + 6, 0.0, -(im * nx.bdy)); put(r + 3, r + 9, 0.0, -(im * nx.bl));
                    }
                    put(r + 2, r, 0.0, j20); put(r + 2, r + 1, 0.0, j21); put(r + 2, r + 4, 0.0, im * me.al);
                    put(r + 3, r, 0.0, j30); put(r + 3, r + 1, 0.0, j31); put(r + 3, r + 4, 0.0, im * me.bl);
                    put(r + 4, r, 0.0, 2 * me.dx); put(r + 4, r + 1, 0.0, 2 * me.dy);
                    if (i > 0) {
                        put(r + 2, r - 5, 0.0, -(im * me.adx)); put(r + 2, r - 4, 0.0, -(im * me.ady));
                        put(r + 3, r - 5, 0.0, -(im * me.bdx)); put(r + 3, r - 4, 0.0, -(im * me.bdy));
                        put(r + 4, r - 5, 0.0, -2 * me.dx); put(r + 4, r - 4, 0.0, -2 * me.dy);
                    }
                }
                lu_band_real(lur, pivr, N);
                lu_band_cplx(lcr, lci, pivc, N);
                cnt[CN_NLU]++;
                lu_valid = true;
                state = S_NEWTON;
            }
        } else if (run == S_NEWTON) {
            if (state == S_NEWTON) {          // one simplified-Newton iteration (:846-940)
                bool finite = true;
#pragma unroll 1
                for (int i = 0; i < 3; ++i) {
                    const double t0c = kStageT[i][0], t1c = kStageT[i][1], t2c = kStageT[i][2];
                    const double a0 = kStageTI[0][i], a1 = kStageTI[1][i], a2 = kStageTI[2][i];
                    rope_rhs(kc,
                             [&](int c) {
                                 const double z = (i == 2) ? (W[c] + W[N + c]) : fma(t2c, W[2 * N + c], fma(t1c, W[N + c], t0c * W[c]));
                                 return y[c] + z;
                             },
                             [&](int c, double Fv) {
                                 finite = finite && is_finite(Fv);
                                 if (i == 0) { R[c] = a0 * Fv; R1[c] = a1 * Fv; R2[c] = a2 * Fv; }
                                 else { R[c] = fma(a0, Fv, R[c]); R1[c] = fma(a1, Fv, R1[c]); R2[c] = fma(a2, Fv, R2[c]); }
                             });
                }
                cnt[CN_NFEV] += 3;
                int outcome = 0;
                bool update_old = true;
                if (!finite) outcome = 2;
                else {
                    for (int c = 0; c < N; ++c) {
                        const double hsc = (a.scale_residuals && vidx[c] >= 1) ? inv_h_lu : 1.0;
                        double ar = R[c], br = R1[c], bi = R2[c];
                        if (c % 5 != 4) {
                            const double w0 = W[c], w1 = W[N + c], w2 = W[2 * N + c];
                            ar = fma(-mr, w0, ar);
                            br = fma(-mcr, w1, br); br = fma(mci, w2, br);
                            bi = fma(-mcr, w2, bi); bi = fma(-mci, w1, bi);
                        }
                        R[c] = ar * hsc; R1[c] = br * hsc; R2[c] = bi * hsc;
                    }
                    solve_band_real(lur, pivr, R, N);
                    solve_band_cplx(lcr, lci, pivc, R1, R2, N);
                    cnt[CN_NSOLVE]++;
                    const double dwn = sqrt(sumsq_scaled(R, inv_ns, N, 3 * N)) * a.rsq3n;
                    for (int e = 0; e < 3 * N; ++e) W[e] += R[e];
                    double dw_true = 0, dw_true_max = 0;
                    if (have_old) {
                        rate = dwn / dwn_old; have_rate = true;
                        const double inv1 = 1.0 / (1.0 - rate);
                        dw_true = rate * inv1 * dwn;
                        double rp = rate;
                        for (int e = 1; e < MAXIT; ++e) rp = (e < MAXIT - k) ? rp * rate : rp;
                        dw_true_max = rp * inv1 * dwn;
                    }
                    if (dwn < a.newton_tol) outcome = 1;
                    else if (have_rate) {
                        bool fall = true;
                        if (rate >= 1) {
                            if (rate < 100 && nbad < a.nmax_bad) { nbad++; update_old = false; fall = false; }
                            else if (!a.all_newton) { outcome = 2; fall = false; }
                        }
                        if (fall) {
                            if (dw_true < a.newton_tol) outcome = 1;
                            else if (dw_true_max > a.newton_tol && a.predictive_newton) {
                                if (nbad < a.nmax_bad) { nbad++; update_old = false; }
                                else if (!a.all_newton) outcome = 2;
                            }
                        }
                    }
                    if (outcome == 0 && update_old) { dwn_old = dwn; have_old = true; }
                }
                if (outcome == 0) {
                    ++k;
                    if (k >= MAXIT) { outcome = 2; n_iter = MAXIT; }
                } else n_iter = k + 1;
                if (outcome != 0) {
                    safety = a.safety_factor * (2 * MAXIT + 1) / (2 * MAXIT + n_iter);
                    if (outcome == 1) state = S_ERREST;
                    else if (current_jac) {
                        cnt[CN_NFAIL]++;
                        h_abs *= a.nonconv_factor;
                        lu_valid = false;
                        state = S_SETUP; setup_kind = SETUP_ATTEMPT;
                    } else {
                        for (int c = 0; c < N; ++c) yJ[c] = y[c];
                        cnt[CN_NJEV]++;
                        current_jac = true; lu_valid = false;
                        state = S_SETUP; setup_kind = SETUP_GUESS;
                    }
                }
            }
        } else if (run == S_ERREST) {
            if (state == S_ERREST) {          // error estimate, accept/reject, epilogue (:660-784)
                // stage values Z = T W into the residual buffer (Newton is over)
                for (int c = 0; c < N; ++c) {
                    const double w0 = W[c], w1 = W[N + c], w2 = W[2 * N + c];
                    R[c] = stage_z<0>(w0, w1, w2); R1[c] = stage_z<1>(w0, w1, w2); R2[c] = stage_z<2>(w0, w1, w2);
                }
                double error_norm = 0.0;
                bool accepted = true;
                if (!a.constant_dt) {
                    for (int c = 0; c < N; ++c) {
                        double v = R[c] * RE0;
                        v = fma(R1[c], RE1, v);
                        v = fma(R2[c], RE2, v);
                        const double zev = v * inv_h;
                        const double yc = y[c];
                        const double ynew = yc + R2[c];
                        const int ve = vidx[c] > 1 ? vidx[c] - 1 : 0;
                        const double hp = a.scale_error ? hpow(h, ve) : 1.0;
                        inv_es[c] = hp / (atol[c] + fmax(fabs(yc), fabs(ynew)) * a.rtol);
                        const double mz = (c % 5 != 4) ? zev : 0.0;
                        ze[c] = mz;
                        err[c] = f[c] + mz;
                    }
                    for (int pass = 0;; ++pass) {
                        solve_band_real(lur, pivr, err, N);
                        cnt[CN_NSOLVE_ERR]++;
                        if (a.zero_algebraic_error)
                            for (int c = 0; c < N; ++c) if (vidx[c] != 0) err[c] = 0.0;
                        error_norm = sqrt(sumsq_scaled(err, inv_es, N, N)) * (a.zero_algebraic_error ? a.rsq_nondiff : a.rsqn);
                        if (pass == 1 || !(error_norm > 1 && (a.always_2nd || rejected))) break;
                        // err <- rhs(t, y + err) + M.ZE, in place: the walk goes from the last node to the first,
                        // reads nodes i and i-1 before it writes node i, and never reads node i again
                        rope_rhs(kc, [&](int c) { return y[c] + err[c]; }, [&](int c, double Fv) { err[c] = Fv + ze[c]; });
                        cnt[CN_NFEV]++;
                    }
                    if (error_norm > 1) {
                        const double factor = gustafsson(h_abs, h_abs_old, error_norm, err_old, hist, a.predictive_controller);
                        h_abs *= fmax(a.min_factor, safety * factor);
                        lu_valid = false; rejected = true; cnt[CN_NREJ]++;
                        accepted = false;
                        state = S_SETUP; setup_kind = SETUP_ATTEMPT;
                    } else cnt[CN_NACC]++;
                }
                if (accepted) {
                    const bool recompute = n_iter > 2 && have_rate && rate > a.jac_recompute;
                    double factor;
                    if (a.constant_dt) factor = a.max_step / h_abs;
                    else factor = fmin(a.max_factor, safety * gustafsson(h_abs, h_abs_old, error_norm, err_old, hist, a.predictive_controller));
                    if (!recompute && factor < 1.2) factor = 1; else lu_valid = false;
                    for (int c = 0; c < N; ++c) {
                        const double z0 = R[c], z1 = R1[c], z2 = R2[c];
                        const double yo = y[c];
                        yold[c] = yo;
                        y[c] = yo + z2;
                        Q[3 * c] = fma(z2, P20, fma(z1, P10, z0 * P00));
                        Q[3 * c + 1] = fma(z2, P21, fma(z1, P11, z0 * P01));
                        Q[3 * c + 2] = fma(z2, P22, fma(z1, P12, z0 * P02));
                    }
                    rope_rhs(kc, [&](int c) { return y[c]; }, [&](int c, double v) { f[c] = v; });
                    cnt[CN_NFEV]++;
                    if (recompute) {
                        for (int c = 0; c < N; ++c) yJ[c] = y[c];
                        cnt[CN_NJEV]++;
                        current_jac = true;
                    } else current_jac = false;
                    h_abs_old_state = h_abs_state; err_old_state = error_norm; hist_state = true;
                    h_abs_state = h_abs * factor;
                    t_sol_old = t; inv_h_sol = inv_h; have_sol = true;
                    t = t_new;
                    while (ei < a.T) {
                        const double te = a.t_eval[ei];
                        if (!(dir > 0 ? te <= t : te >= t)) break;
                        const double x = (te - t_sol_old) * inv_h_sol;
                        const double x2 = x * x, x3 = x2 * x;
                        for (int c = 0; c < N; ++c) {
                            double v = Q[3 * c] * x;
                            v = fma(Q[3 * c + 1], x2, v);
                            v = fma(Q[3 * c + 2], x3, v);
                            a.y_eval[((long long)ei * N + c) * a.B + idx] = v + yold[c];
                        }
                        ++ei;
                    }
                    if (dir * (t - a.tb) >= 0) finish_status = BRDAE_STATUS_FINISHED;
                    else { state = S_SETUP; setup_kind = SETUP_STEP; }
                }
            }
        }
        __syncwarp();
        // ---------------- cheap set-up: prologue, attempt, starting guess
        if (__any_sync(0xffffffffu, state == S_SETUP && finish_status == 1)) {
            if (state == S_SETUP && finish_status == 1) {
                if (setup_kind == SETUP_STEP) {
                    ++nsteps;
                    if (a.nmax_step > 0 && nsteps > a.nmax_step) finish_status = BRDAE_STATUS_NMAX_STEP;
                    else {
                        cnt[CN_NSTEP]++;
                        min_step = fmax(1e-20, 10 * fabs(nextafter(t, dir * INFINITY) - t));
                        h_abs_old = h_abs_old_state; err_old = err_old_state; hist = hist_state;
                        if (h_abs_state > a.max_step) { h_abs = a.max_step; hist = false; }
                        else if (h_abs_state < min_step) { h_abs = min_step; hist = false; }
                        else h_abs = h_abs_state;
                        if (fabs(a.tb - (t + dir * h_abs)) < 1e-2 * h_abs) { h_abs = fabs(a.tb - t); lu_valid = false; }
                        rejected = false;
                    }
                }
                if (finish_status == 1 && setup_kind <= SETUP_ATTEMPT) {
                    if (h_abs < min_step) finish_status = BRDAE_STATUS_TOO_SMALL_STEP;
                    else {
                        h = h_abs * dir;
                        t_new = t + h;
                        if (dir * (t_new - a.tb) > 0) t_new = a.tb;
                        h = t_new - t;
                        h_abs = fabs(h);
                        inv_h = 1.0 / h;
                        for (int c = 0; c < N; ++c) {
                            const int ve = vidx[c] > 1 ? vidx[c] - 1 : 0;
                            const double hp = a.scale_newton_norm ? hpow(h, ve) : 1.0;
                            inv_ns[c] = hp / (atol[c] + fabs(y[c]) * a.rtol);
                        }
                        mr = MU_REAL * inv_h; mcr = MU_CR * inv_h; mci = MU_CI * inv_h;
                    }
                }
                if (finish_status == 1) {
                    if (!have_sol || !a.extrapolate) {
                        for (int e = 0; e < 3 * N; ++e) W[e] = 0.0;
                    } else {
                        double x[3], x2[3], x3[3];
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const double tt = t + h * stage_c(i);
                            x[i] = (tt - t_sol_old) * inv_h_sol; x2[i] = x[i] * x[i]; x3[i] = x2[i] * x[i];
                        }
                        for (int c = 0; c < N; ++c) {
                            const double q0 = Q[3 * c], q1 = Q[3 * c + 1], q2 = Q[3 * c + 2], yo = yold[c], yc = y[c];
                            double z[3];
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                double v = q0 * x[i];
                                v = fma(q1, x2[i], v);
                                v = fma(q2, x3[i], v);
                                v = v + yo;
                                z[i] = v - yc;
                            }
                            W[c] = fma(TI02, z[2], fma(TI01, z[1], TI00 * z[0]));
                            W[N + c] = fma(TI12, z[2], fma(TI11, z[1], TI10 * z[0]));
                            W[2 * N + c] = fma(TI22, z[2], fma(TI21, z[1], TI20 * z[0]));
                        }
                    }
                    k = 0; nbad = 0; have_old = false; have_rate = false; rate = 0; dwn_old = 0;
                    state = lu_valid ? S_NEWTON : S_FACTOR;
                }
            }
        }
        // ---------------- retire
        if (__any_sync(0xffffffffu, finish_status != 1)) {
            if (finish_status != 1) {
                for (int e = ei; e < a.T; ++e)
                    for (int c = 0; c < N; ++c) a.y_eval[((long long)e * N + c) * a.B + idx] = NAN;
                for (int c = 0; c < N; ++c) a.y_final[(long long)c * a.B + idx] = y[c];
                a.n_eval[idx] = ei; a.t_final[idx] = t; a.status[idx] = finish_status;
#pragma unroll
                for (int q = 0; q < CN_COUNT; ++q) a.counters[(long long)q * a.B + idx] = cnt[q];
                state = S_IDLE;
            }
        }
    }
}

constexpr size_t kHeaderBytes = 256;
size_t align8(size_t x) { return (x + 7) / 8 * 8; }

}  // namespace

// number of lane slots used for B ropes: one per rope up to a memory cap of about 24 GB of state
long long k3_slots(int n, int64_t B) {
    const long long per_slot = 95LL * n * 8 + 2LL * n * 4;
    long long cap = (24LL << 30) / per_slot;
    cap = cap / K3_BLOCK * K3_BLOCK;
    if (cap < K3_BLOCK) cap = K3_BLOCK;
    long long G = (B + K3_BLOCK - 1) / K3_BLOCK * K3_BLOCK;
    return G < cap ? G : cap;
}

// workspace after the 256-byte work counter: atol[N] | var_index[N] | state [95 N][G] | ipiv [2 N][G]
size_t k3_scratch_bytes(int n, int64_t B) {
    const long long G = k3_slots(n, B);
    return align8(sizeof(double) * n) + align8(sizeof(int32_t) * n) + (size_t)95 * n * G * sizeof(double) +
           (size_t)2 * n * G * sizeof(int);
}

int k3_npendulum(bool dense_mass, const SolveArgs &a_in, const brdae_batch *b, int sms, cudaStream_t stream, LaunchGeom *g,
                 bool geom_only) {
    (void)sms;
    const int N = a_in.n;
    if (N < 10 || N % 5 != 0) return BRDAE_ERR_BAD_ARGUMENT;
    const long long G = k3_slots(N, a_in.B);
    if (g) { g->grid = (int)(G / K3_BLOCK); g->block = K3_BLOCK; g->smem = 0; g->scratch_bytes = 0; }
    if (geom_only) return BRDAE_OK;
    if (dense_mass || a_in.has_jac_const) return BRDAE_ERR_UNSUPPORTED;
    SolveArgs a = a_in;
    char *ws = reinterpret_cast<char *>(a.work_counter) + kHeaderBytes;
    double *atol_d = reinterpret_cast<double *>(ws);
    int32_t *vidx_d = reinterpret_cast<int32_t *>(ws + align8(sizeof(double) * N));
    char *after = ws + align8(sizeof(double) * N) + align8(sizeof(int32_t) * N);
    if (cudaMemcpyAsync(atol_d, b->atol, sizeof(double) * N, cudaMemcpyHostToDevice, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    if (b->var_index) {
        if (cudaMemcpyAsync(vidx_d, b->var_index, sizeof(int32_t) * N, cudaMemcpyHostToDevice, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    } else if (cudaMemsetAsync(vidx_d, 0, sizeof(int32_t) * N, stream) != cudaSuccess) return BRDAE_ERR_CUDA;
    a.atol_dev = atol_d; a.var_index_dev = vidx_d;
    K3Cfg cfg;
    cfg.state = reinterpret_cast<double *>(after);
    cfg.ipiv = reinterpret_cast<int *>(after + (size_t)95 * N * G * sizeof(double));
    cfg.G = G;
    k3_kernel<<<(int)(G / K3_BLOCK), K3_BLOCK, 0, stream>>>(a, cfg);
    return cudaPeekAtLastError() == cudaSuccess ? BRDAE_OK : BRDAE_ERR_CUDA;
}

}  // namespace brdae
