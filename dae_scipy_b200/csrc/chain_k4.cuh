// chain_k4.cuh -- K4: lane-per-rope batched RadauDAE for the recursive n-pendulum chain
// (reference problem scipyDAE/tests/recursive_pendulum.py:93-127; N = 5 n unknowns interleaved
// (x, y, vx, vy, lambda) per node, index 3, mass diag(1,1,1,1,0) per node), sm_100a.
//
// The band kernel K2 factorises the 5n x 5n iteration matrices as general banded matrices: N sequential
// pivoted column steps on one warp, 2 N kl (kl+ku) flops, 1-1.6 % of the FP64 peak (profiles/r01_k2_ncu_summary.md).
// This kernel uses what the chain IS: mu M - J is block tridiagonal in the nodes, the rows x_i, y_i read
// mu p - v = r, so the velocities are eliminated exactly and a 3 x 3 block in (dx, dy, dlambda) per node is left.
// The blocks are eliminated node by node (block Thomas), each pivot block E_i inverted explicitly with partial
// pivoting inside the block ("chain spec"; the scalar restatement used by the tests states the same operations in the same order):
// 14 x fewer flops than the band LU, n sequential block steps instead of 5 n column steps, and no step needs more
// than one lane.  So ONE LANE integrates one rope (K1's control structure), with the rope's state -- 99 doubles per
// node: y, f, y_old, Q, W, the rod quantities of the Jacobian, the Newton scales, the three inverse pivot blocks, the
// forward-substituted blocks -- in global memory, laid out [warp][node][slot][lane] so that every access of a warp is one
// coalesced 256-byte line and a sweep over the nodes streams through one contiguous region.  A Newton iteration is two
// sweeps: forward (stage right-hand sides with a one-node look-ahead, residuals, velocity elimination, block forward
// substitution) and backward (block back substitution, velocities, norm, W += dW).  The bound is HBM / L2 bandwidth
// (about 1.2 KB per node and Newton iteration), not the FP64 pipe.
//
// Reference map: radauDAE.py:515-791 (step), 793-949 (Newton), 951-982 (dense output), 1079-1157 (driver / t_eval).
#pragma once
#include "radau_k1.cuh"     // warp helpers (warp_any, fetch_system, ...), state names

namespace brdae {
namespace k4 {

constexpr double GRAV = 9.81;
// Slots of one node, ordered so that what a sweep READS is one contiguous range (the range is what the sweep stages in shared
// memory, see Stager): set-up reads [O_Q, O_W), a Newton forward sweep [O_Y, O_Z), a Newton backward sweep [O_W, SPN), ...
enum : int {
    O_Q = 0 /* [c][3] */, O_YO = 15, O_F = 20, O_Y = 25, O_W = 30 /* [q][c] */, O_ND = 45 /* adx ady bdy al bl dx2 dy2 */,
    O_VR = 52, O_VCR = 61, O_VCI = 70, O_Z = 79 /* zr[3] zcr[3] zci[3] */, O_RP = 88 /* r[2] cr[2] ci[2] */, O_NS = 94,
    SPN = 99,
    // error-estimate phase: the complex slots are free
    O_ERR = O_Z + 3 /* [5] */, O_MZE = O_RP + 2 /* [4] */,
    STAGE_SLOTS = SPN - O_W     // the widest range any sweep stages (Newton backward: 69 slots)
};
enum : int { ND_ADX = 0, ND_ADY, ND_BDY, ND_AL, ND_BL, ND_DX2, ND_DY2 };

struct Consts { int nn; double m_in, im_in, r0s; };
BRDAE_HD double node_m(const Consts &k, int i) { return (i == k.nn - 1) ? 1.0 : k.m_in; }
BRDAE_HD double node_im(const Consts &k, int i) { return (i == k.nn - 1) ? 1.0 : k.im_in; }

template <int STRIDE>
struct Rope {
    double *p;
    BRDAE_HD double *node(int i) const { return p + (size_t)i * (SPN * STRIDE); }
};
// A node as a sweep sees it: loads come from `ld` (the staged copy in shared memory, addressed with the same slot offsets, or
// the node itself), stores go to `st` (always the node in global memory).
struct NodeRef { const double *ld; double *st; };
#define K4L(nb, s) (nb).ld[(s) * STRIDE]
#define K4W(nb, s) (nb).st[(s) * STRIDE]
#define K4G(ptr, s) (ptr)[(s) * STRIDE]      // plain global access (loading a rope, output, retirement)

// ---- staging of the node stream through shared memory (TMA bulk copies) ------------------------------------------------------
// A sweep walks the nodes of a rope in order and reads 20-69 slots of each; the ropes of a warp keep theirs side by side
// ([node][slot][rope]), so what the warp reads of one node is ONE contiguous block of global memory.  The first version read it
// with ordinary loads: 16,384 ropes stream 0.7 TB through HBM per launch (ncu: 2.3 TB/s, 61 % of the warp cycles stalled on the
// long scoreboard, 252 registers -- none left to issue the next node's loads early; L1/L2 prefetch instructions gained 6 %,
// profiles/r02_k4_ncu_summary.md).  Here the warp's elected lane asks the TMA unit for the block of the node(s) ahead
// (cp.async.bulk global -> shared, completion on an mbarrier) while the lanes compute the current node out of shared memory:
// a ring of three node buffers per warp -- current and two ahead (forward sweeps read the next node's state during a visit; backward
// sweeps carry what they need of the previous node in registers).  Stores go straight to global memory; a sweep begins with
// fence.proxy.async so that the stores of the sweeps before it are ordered before its bulk reads.
// Buffer b's k-th use completes phase k of its mbarrier; the use counts live next to the barriers in shared memory because the
// set of active lanes (and with it the elected lane) changes from sweep to sweep.
#ifndef BRDAE_K4_UNROLL
#define BRDAE_K4_UNROLL 1        // node loops of the Newton sweeps: 1 = rolled
#endif
#define K4_STR2(x) #x
#define K4_STR(x) K4_STR2(x)
#define K4_NEWTON_UNROLL K4_STR(unroll BRDAE_K4_UNROLL)
#ifndef BRDAE_K4_STAGED
#define BRDAE_K4_STAGED 1
#endif
template <int STRIDE, bool STAGED>
struct Stager;

template <int STRIDE>
struct Stager<STRIDE, false> {        // no staging: the node itself (the host harness, one-rope warps, A/B builds)
    const Rope<STRIDE> *R;
    BRDAE_HD void begin(const Rope<STRIDE> &R_, int, int, int, int, int, unsigned) { R = &R_; }
    BRDAE_HD void step(int) {}
    BRDAE_HD NodeRef ref(int i) const { double *q = R->node(i); return NodeRef{q, q}; }
};

#if defined(__CUDA_ARCH__)
template <int STRIDE>
struct Stager<STRIDE, true> {
    static constexpr int BUF_BYTES = STAGE_SLOTS * STRIDE * 8;
    static constexpr int RING_BYTES = 3 * BUF_BYTES;
    static constexpr int SMEM_BYTES = RING_BYTES + 64;           // ring | 3 mbarriers (8 B each) | 3 use counts (4 B each)
    unsigned ring;              // shared-memory address of the ring (this warp's dynamic shared memory, 128-byte aligned)
    const double *ring_g;       // the same as a generic pointer
    const Rope<STRIDE> *R;
    const char *gsrc;           // global address of slot `lo` of node 0, rope 0 of the warp
    unsigned mask;
    int lane, lo, bytes, dir, nn, next, waited;
    bool leader;

    __device__ void init(void *smem) {          // once per kernel, all 32 lanes
        ring_g = reinterpret_cast<const double *>(smem);
        ring = (unsigned)__cvta_generic_to_shared(smem);
        lane = threadIdx.x & 31;
        if (lane == 0) {
            for (int b = 0; b < 3; ++b) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ring + RING_BYTES + 8 * b));
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(ring + RING_BYTES + 24 + 4 * b), "r"(0));
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
    __device__ unsigned bar(int b) const { return ring + RING_BYTES + 8 * b; }
    __device__ unsigned uses(int b) const { return ring + RING_BYTES + 24 + 4 * b; }
    __device__ void issue(int j) {              // elected lane: node j -> buffer j % 3
        const int b = j % 3;
        const unsigned dst = ring + b * BUF_BYTES, mb = bar(b);
        const char *src = gsrc + (size_t)j * (SPN * STRIDE * 8);
        unsigned u;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(uses(b)));
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(uses(b)), "r"(u + 1u));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
    }
    __device__ void wait(int j) {               // every active lane: node j has arrived
        const int b = j % 3;
        unsigned u;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(uses(b)));
        const unsigned parity = (u - 1u) & 1u;
        unsigned done;
        do {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(bar(b)), "r"(parity) : "memory");
        } while (!done);
    }
    // nodes first, first + dir_, ... of the rope are visited by the lanes of `m` (a ballot of the whole warp taken right before
    // the sweep); slots [lo_, lo_ + cnt) of each node are staged
    __device__ void begin(const Rope<STRIDE> &R_, int lo_, int cnt, int first, int dir_, int nn_, unsigned m) {
        R = &R_; lo = lo_; bytes = cnt * STRIDE * 8; dir = dir_; nn = nn_;
        mask = m;
        leader = lane == __ffs(mask) - 1;
        gsrc = reinterpret_cast<const char *>(R_.p - lane) + (size_t)lo * (STRIDE * 8);
        asm volatile("fence.proxy.async.global;" ::: "memory");     // this lane's earlier stores before the bulk reads below
        __syncwarp(mask);
        next = first; waited = first - dir_;
        fill(first + dir_);          // (the third buffer is requested by the first step())
    }
    __device__ void fill(int upto) {            // issue every node up to `upto` (inclusive) that has not been issued yet
        while (dir > 0 ? (next <= upto && next < nn) : (next >= upto && next >= 0)) {
            if (leader) issue(next);
            next += dir;
        }
    }
    // top of the visit of node i: the buffer of the node visited before is free (what backward sweeps still need of it, its
    // rod quantities, they carry in registers): two nodes ahead are in flight while this one is computed; forward sweeps read
    // the next node during this visit (look-ahead), so that one must have arrived
    __device__ void step(int i) {
        __syncwarp(mask);
        fill(dir > 0 ? i + 2 : i - 2);
        __syncwarp(mask);                       // the use counts written by the elected lane
        const int need = dir > 0 ? (i + 1 < nn ? i + 1 : i) : i;
        while (waited != need) { waited += dir; wait(waited); }
    }
    __device__ NodeRef ref(int i) const {
        return NodeRef{ring_g + ((i % 3) * (STAGE_SLOTS * STRIDE) - lo * STRIDE + lane), R->node(i)};
    }
};
#endif

struct Rod { double a, b, rs; };
// lambda dx / rs, lambda dy / rs with one reciprocal per rod (oracle npe_rhs)
BRDAE_HD Rod rod(double dx, double dy, double l) {
    Rod r;
    r.rs = dx * dx + dy * dy;
    const double ir = 1.0 / r.rs;
    r.a = (l * dx) * ir;
    r.b = (l * dy) * ir;
    return r;
}

// f = rhs(Y) (recursive_pendulum.py:93-118), walked forward with a one-node look-ahead; Yf(i, c) returns the state,
// Ff(i, c, v) consumes f, Nf(i, dx, dy, l) (optional) sees the rod of node i
template <class YF, class FF, class NF>
BRDAE_HD void rhs_sweep(const Consts &k, YF &&Yf, FF &&Ff, NF &&Nf) {
    double x = Yf(0, 0), y = Yf(0, 1), l = Yf(0, 4);
    double dx = x, dy = y;
    Rod cur = rod(dx, dy, l);
    for (int i = 0; i < k.nn; ++i) {
        const bool last = (i == k.nn - 1);
        Rod nxt = cur;
        double xn = 0, yn = 0, ln = 0, dxn = 0, dyn = 0;
        if (!last) {
            xn = Yf(i + 1, 0); yn = Yf(i + 1, 1); ln = Yf(i + 1, 4);
            dxn = xn - x; dyn = yn - y;
            nxt = rod(dxn, dyn, ln);
        }
        Nf(i, dx, dy, l);
        const double vx = Yf(i, 2), vy = Yf(i, 3);
        const double m = node_m(k, i), im = node_im(k, i);
        Ff(i, 0, vx);
        Ff(i, 1, vy);
        if (last) { Ff(i, 2, im * cur.a); Ff(i, 3, im * (cur.b - m * GRAV)); }
        else { Ff(i, 2, im * (cur.a - nxt.a)); Ff(i, 3, im * (cur.b - nxt.b - m * GRAV)); }
        Ff(i, 4, cur.rs - k.r0s);
        x = xn; y = yn; l = ln; dx = dxn; dy = dyn; cur = nxt;
    }
}
// rod quantities of the analytic Jacobian (oracle chain_jac)
template <int STRIDE>
BRDAE_HD void store_node_jac(double *st, double dx, double dy, double l) {
    const double rs = dx * dx + dy * dy, rs2 = rs * rs;
    K4G(st, O_ND + ND_ADX) = l * (rs - 2 * dx * dx) / rs2;
    K4G(st, O_ND + ND_ADY) = l * (-2 * dx * dy) / rs2;
    K4G(st, O_ND + ND_BDY) = l * (rs - 2 * dy * dy) / rs2;
    K4G(st, O_ND + ND_AL) = dx / rs;
    K4G(st, O_ND + ND_BL) = dy / rs;
    K4G(st, O_ND + ND_DX2) = 2 * dx;
    K4G(st, O_ND + ND_DY2) = 2 * dy;
}
template <int STRIDE, class STG>
BRDAE_HD void jac_sweep(STG &S, const Rope<STRIDE> &R, const Consts &k, unsigned m) {
    double xp = 0, yp = 0;
    S.begin(R, O_Y, 5, 0, 1, k.nn, m);
    for (int i = 0; i < k.nn; ++i) {
        S.step(i);
        const NodeRef nb = S.ref(i);
        const double x = K4L(nb, O_Y), y = K4L(nb, O_Y + 1), l = K4L(nb, O_Y + 4);
        const double dx = (i == 0) ? x : x - xp, dy = (i == 0) ? y : y - yp;
        store_node_jac<STRIDE>(nb.st, dx, dy, l);
        xp = x; yp = y;
    }
}

// ---- 3 x 3 inverses, Gauss-Jordan with partial pivoting (oracle inv3_real / inv3_cplx), in registers ----
BRDAE_HD void inv3_real(double (&E)[3][3], double (&V)[3][3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
        int p = kk;
        double best = fabs(E[kk][kk]);
#pragma unroll
        for (int r = kk + 1; r < 3; ++r) { const double v = fabs(E[r][kk]); const bool g = v > best; best = g ? v : best; p = g ? r : p; }
#pragma unroll
        for (int r = kk + 1; r < 3; ++r) {          // exchange rows kk and p (selects: no dynamic register index)
            const bool sw = (p == r);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double ek = E[kk][c], er = E[r][c], vk = V[kk][c], vr = V[r][c];
                E[kk][c] = sw ? er : ek; E[r][c] = sw ? ek : er;
                V[kk][c] = sw ? vr : vk; V[r][c] = sw ? vk : vr;
            }
        }
        const double rcp = 1.0 / E[kk][kk];
#pragma unroll
        for (int c = kk + 1; c < 3; ++c) E[kk][c] = E[kk][c] * rcp;
#pragma unroll
        for (int c = 0; c < 3; ++c) V[kk][c] = V[kk][c] * rcp;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (r == kk) continue;
            const double f = E[r][kk];
#pragma unroll
            for (int c = kk + 1; c < 3; ++c) E[r][c] = fma(-f, E[kk][c], E[r][c]);
#pragma unroll
            for (int c = 0; c < 3; ++c) V[r][c] = fma(-f, V[kk][c], V[r][c]);
        }
    }
}
BRDAE_HD void inv3_cplx(double (&Er)[3][3], double (&Ei)[3][3], double (&Vr)[3][3], double (&Vi)[3][3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) { Vr[r][c] = (r == c) ? 1.0 : 0.0; Vi[r][c] = 0.0; }
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
        int p = kk;
        double best = fabs(Er[kk][kk]) + fabs(Ei[kk][kk]);
#pragma unroll
        for (int r = kk + 1; r < 3; ++r) {
            const double v = fabs(Er[r][kk]) + fabs(Ei[r][kk]);
            const bool g = v > best;
            best = g ? v : best; p = g ? r : p;
        }
#pragma unroll
        for (int r = kk + 1; r < 3; ++r) {
            const bool sw = (p == r);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double a = Er[kk][c], b = Er[r][c];
                Er[kk][c] = sw ? b : a; Er[r][c] = sw ? a : b;
                a = Ei[kk][c]; b = Ei[r][c];
                Ei[kk][c] = sw ? b : a; Ei[r][c] = sw ? a : b;
                a = Vr[kk][c]; b = Vr[r][c];
                Vr[kk][c] = sw ? b : a; Vr[r][c] = sw ? a : b;
                a = Vi[kk][c]; b = Vi[r][c];
                Vi[kk][c] = sw ? b : a; Vi[r][c] = sw ? a : b;
            }
        }
        const double pr = Er[kk][kk], pi = Ei[kk][kk];
        const double invd = 1.0 / fma(pr, pr, pi * pi);
        const double qr = pr * invd, qi = -(pi * invd);
#pragma unroll
        for (int c = kk + 1; c < 3; ++c) {
            const double xr = Er[kk][c], xi = Ei[kk][c];
            Er[kk][c] = fma(xr, qr, -(xi * qi)); Ei[kk][c] = fma(xr, qi, xi * qr);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double xr = Vr[kk][c], xi = Vi[kk][c];
            Vr[kk][c] = fma(xr, qr, -(xi * qi)); Vi[kk][c] = fma(xr, qi, xi * qr);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (r == kk) continue;
            const double fr = Er[r][kk], fi = Ei[r][kk];
#pragma unroll
            for (int c = kk + 1; c < 3; ++c) {
                const double ur = Er[kk][c], ui = Ei[kk][c];
                double xr = Er[r][c], xi = Ei[r][c];
                xr = fma(-fr, ur, xr); xr = fma(fi, ui, xr);
                xi = fma(-fr, ui, xi); xi = fma(-fi, ur, xi);
                Er[r][c] = xr; Ei[r][c] = xi;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double ur = Vr[kk][c], ui = Vi[kk][c];
                double xr = Vr[r][c], xi = Vi[r][c];
                xr = fma(-fr, ur, xr); xr = fma(fi, ui, xr);
                xi = fma(-fr, ui, xi); xi = fma(-fi, ur, xi);
                Vr[r][c] = xr; Vi[r][c] = xi;
            }
        }
    }
}

// ---- both factorisations (oracle chain_factor): E_0 = D_0, E_i = D_i - L_i E_{i-1}^-1 U_{i-1}, V_i = E_i^-1 ----
template <int STRIDE, class STG>
BRDAE_HD void factor_sweep(STG &S, const Rope<STRIDE> &R, const Consts &k, double mr, double mcr, double mci, unsigned m) {
    const double m2 = mr * mr;
    const double m2r = fma(mcr, mcr, -(mci * mci)), m2i = 2 * (mcr * mci);
    double G[2][2] = {{0, 0}, {0, 0}}, Gr[2][2] = {{0, 0}, {0, 0}}, Gi[2][2] = {{0, 0}, {0, 0}};
    S.begin(R, O_ND, 7, 0, 1, k.nn, m);
    S.step(0);
    double adx, ady, bdy, al, bl, dx2, dy2;
    {
        const NodeRef n0 = S.ref(0);
        adx = K4L(n0, O_ND + ND_ADX); ady = K4L(n0, O_ND + ND_ADY); bdy = K4L(n0, O_ND + ND_BDY);
        al = K4L(n0, O_ND + ND_AL); bl = K4L(n0, O_ND + ND_BL); dx2 = K4L(n0, O_ND + ND_DX2); dy2 = K4L(n0, O_ND + ND_DY2);
    }
    for (int i = 0; i < k.nn; ++i) {
        S.step(i);
        const NodeRef nb = S.ref(i);
        const double im = node_im(k, i);
        double sxx = im * adx, sxy = im * ady, syy = im * bdy;
        double adxn = 0, adyn = 0, bdyn = 0, aln = 0, bln = 0, dx2n = 0, dy2n = 0;
        if (i < k.nn - 1) {
            const NodeRef nn_ = S.ref(i + 1);
            adxn = K4L(nn_, O_ND + ND_ADX); adyn = K4L(nn_, O_ND + ND_ADY); bdyn = K4L(nn_, O_ND + ND_BDY);
            aln = K4L(nn_, O_ND + ND_AL); bln = K4L(nn_, O_ND + ND_BL); dx2n = K4L(nn_, O_ND + ND_DX2); dy2n = K4L(nn_, O_ND + ND_DY2);
            sxx = sxx + im * adxn; sxy = sxy + im * adyn; syy = syy + im * bdyn;
        }
        const double e02 = -(im * al), e12 = -(im * bl);
        double Er[3][3] = {{m2 - sxx, -sxy, e02}, {-sxy, m2 - syy, e12}, {-dx2, -dy2, 0.0}};
        double Cr[3][3] = {{m2r - sxx, -sxy, e02}, {-sxy, m2r - syy, e12}, {-dx2, -dy2, 0.0}};
        double Ci[3][3] = {{m2i, 0.0, 0.0}, {0.0, m2i, 0.0}, {0.0, 0.0, 0.0}};
        if (i > 0) {
            const double imp = node_im(k, i - 1);
            const double L[3][2] = {{im * adx, im * ady}, {im * ady, im * bdy}, {dx2, dy2}};
            const double U[2][3] = {{imp * adx, imp * ady, imp * al}, {imp * ady, imp * bdy, imp * bl}};
            double X[2][3], Xr[2][3], Xi[2][3];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    X[r][c] = fma(G[r][1], U[1][c], G[r][0] * U[0][c]);
                    Xr[r][c] = fma(Gr[r][1], U[1][c], Gr[r][0] * U[0][c]);
                    Xi[r][c] = fma(Gi[r][1], U[1][c], Gi[r][0] * U[0][c]);
                }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    Er[r][c] = Er[r][c] - fma(L[r][1], X[1][c], L[r][0] * X[0][c]);
                    Cr[r][c] = Cr[r][c] - fma(L[r][1], Xr[1][c], L[r][0] * Xr[0][c]);
                    Ci[r][c] = Ci[r][c] - fma(L[r][1], Xi[1][c], L[r][0] * Xi[0][c]);
                }
        }
        double V[3][3], Vr[3][3], Vi[3][3];
        inv3_real(Er, V);
        inv3_cplx(Cr, Ci, Vr, Vi);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                K4W(nb, O_VR + 3 * r + c) = V[r][c]; K4W(nb, O_VCR + 3 * r + c) = Vr[r][c]; K4W(nb, O_VCI + 3 * r + c) = Vi[r][c];
            }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) { G[r][c] = V[r][c]; Gr[r][c] = Vr[r][c]; Gi[r][c] = Vi[r][c]; }
        adx = adxn; ady = adyn; bdy = bdyn; al = aln; bl = bln; dx2 = dx2n; dy2 = dy2n;
    }
}

// one step of the real forward substitution at node i: e = right-hand side (x, y, vx, vy, lambda rows); stores z and the
// position rows (needed to recover the velocities), updates zp = z[0:2]
template <int STRIDE>
BRDAE_HD void fwd_real(const NodeRef &nb, const Consts &k, int i, double mrf, const double (&e)[5], double (&zp)[2]) {
    double g[3] = {fma(mrf, e[0], e[2]), fma(mrf, e[1], e[3]), e[4]};
    if (i > 0) {
        const double im = node_im(k, i);
        const double adx = K4L(nb, O_ND + ND_ADX), ady = K4L(nb, O_ND + ND_ADY), bdy = K4L(nb, O_ND + ND_BDY);
        const double L[3][2] = {{im * adx, im * ady}, {im * ady, im * bdy}, {K4L(nb, O_ND + ND_DX2), K4L(nb, O_ND + ND_DY2)}};
#pragma unroll
        for (int q = 0; q < 3; ++q) g[q] = fma(-L[q][1], zp[1], fma(-L[q][0], zp[0], g[q]));
    }
    double z[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        z[q] = fma(K4L(nb, O_VR + 3 * q + 2), g[2], fma(K4L(nb, O_VR + 3 * q + 1), g[1], K4L(nb, O_VR + 3 * q) * g[0]));
        K4W(nb, O_Z + q) = z[q];
    }
    K4W(nb, O_RP) = e[0]; K4W(nb, O_RP + 1) = e[1];
    zp[0] = z[0]; zp[1] = z[1];
}
// one step of the real back substitution at node i: u = solution block of node i+1 on entry, of node i on exit;
// d = (dx, dy, dvx, dvy, dlambda) of node i
template <int STRIDE>
// `rn` = the rod quantities (adx, ady, bdy, al, bl) of node i+1, carried in registers from its visit: its buffer is already
// being refilled with the node two visits ahead
BRDAE_HD void bwd_real(const NodeRef &nb, const double (&rn)[5], const Consts &k, int i, double mrf, double (&u)[3], double (&d)[5]) {
    double un[3];
    if (i == k.nn - 1) {
#pragma unroll
        for (int q = 0; q < 3; ++q) un[q] = K4L(nb, O_Z + q);
    } else {
        const double im = node_im(k, i);
        const double adx = rn[0], ady = rn[1], bdy = rn[2];
        const double U[2][3] = {{im * adx, im * ady, im * rn[3]}, {im * ady, im * bdy, im * rn[4]}};
        const double w0 = fma(U[0][2], u[2], fma(U[0][1], u[1], U[0][0] * u[0]));
        const double w1 = fma(U[1][2], u[2], fma(U[1][1], u[1], U[1][0] * u[0]));
#pragma unroll
        for (int q = 0; q < 3; ++q) un[q] = fma(-K4L(nb, O_VR + 3 * q + 1), w1, fma(-K4L(nb, O_VR + 3 * q), w0, K4L(nb, O_Z + q)));
    }
    d[0] = un[0]; d[1] = un[1];
    d[2] = fma(mrf, un[0], -K4L(nb, O_RP));
    d[3] = fma(mrf, un[1], -K4L(nb, O_RP + 1));
    d[4] = un[2];
    u[0] = un[0]; u[1] = un[1]; u[2] = un[2];
}

// ---- one simplified-Newton iteration (:846-940): returns false if a stage right-hand side is not finite; otherwise the
// sum of squares of the scaled increment in `ss` and W += dW done ----
template <int STRIDE, class STG>
BRDAE_HD bool newton_sweeps(STG &S, const Rope<STRIDE> &R, const Consts &k, double mr, double mcr, double mci, double mrf,
                            double mcrf, double mcif, double &ss, unsigned m) {
    bool finite = true;
    {   // ---------- forward: stage right-hand sides, residuals, velocity elimination, block forward substitution
        double cx[3], cy[3];                       // stage positions of the current node
        Rod cur[3];
        double zp[2] = {0, 0}, zpr[2] = {0, 0}, zpi[2] = {0, 0};
        S.begin(R, O_Y, O_Z - O_Y, 0, 1, k.nn, m);
        S.step(0);
        {
            const NodeRef nb = S.ref(0);
            const double y0 = K4L(nb, O_Y), y1 = K4L(nb, O_Y + 1), y4 = K4L(nb, O_Y + 4);
            const double w00 = K4L(nb, O_W), w10 = K4L(nb, O_W + 5), w20 = K4L(nb, O_W + 10);
            const double w01 = K4L(nb, O_W + 1), w11 = K4L(nb, O_W + 6), w21 = K4L(nb, O_W + 11);
            const double w04 = K4L(nb, O_W + 4), w14 = K4L(nb, O_W + 9), w24 = K4L(nb, O_W + 14);
            cx[0] = y0 + stage_z<0>(w00, w10, w20); cx[1] = y0 + stage_z<1>(w00, w10, w20); cx[2] = y0 + stage_z<2>(w00, w10, w20);
            cy[0] = y1 + stage_z<0>(w01, w11, w21); cy[1] = y1 + stage_z<1>(w01, w11, w21); cy[2] = y1 + stage_z<2>(w01, w11, w21);
            const double l0 = y4 + stage_z<0>(w04, w14, w24), l1 = y4 + stage_z<1>(w04, w14, w24), l2 = y4 + stage_z<2>(w04, w14, w24);
            cur[0] = rod(cx[0], cy[0], l0); cur[1] = rod(cx[1], cy[1], l1); cur[2] = rod(cx[2], cy[2], l2);
        }
_Pragma(K4_NEWTON_UNROLL)
        for (int i = 0; i < k.nn; ++i) {
            S.step(i);
            const NodeRef nb = S.ref(i);
            const bool last = (i == k.nn - 1);
            Rod nxt[3] = {cur[0], cur[1], cur[2]};
            double nx[3] = {0, 0, 0}, ny[3] = {0, 0, 0};
            if (!last) {
                const NodeRef nn_ = S.ref(i + 1);
                const double y0 = K4L(nn_, O_Y), y1 = K4L(nn_, O_Y + 1), y4 = K4L(nn_, O_Y + 4);
                const double w00 = K4L(nn_, O_W), w10 = K4L(nn_, O_W + 5), w20 = K4L(nn_, O_W + 10);
                const double w01 = K4L(nn_, O_W + 1), w11 = K4L(nn_, O_W + 6), w21 = K4L(nn_, O_W + 11);
                const double w04 = K4L(nn_, O_W + 4), w14 = K4L(nn_, O_W + 9), w24 = K4L(nn_, O_W + 14);
                nx[0] = y0 + stage_z<0>(w00, w10, w20); nx[1] = y0 + stage_z<1>(w00, w10, w20); nx[2] = y0 + stage_z<2>(w00, w10, w20);
                ny[0] = y1 + stage_z<0>(w01, w11, w21); ny[1] = y1 + stage_z<1>(w01, w11, w21); ny[2] = y1 + stage_z<2>(w01, w11, w21);
                const double l0 = y4 + stage_z<0>(w04, w14, w24), l1 = y4 + stage_z<1>(w04, w14, w24), l2 = y4 + stage_z<2>(w04, w14, w24);
                nxt[0] = rod(nx[0] - cx[0], ny[0] - cy[0], l0);
                nxt[1] = rod(nx[1] - cx[1], ny[1] - cy[1], l1);
                nxt[2] = rod(nx[2] - cx[2], ny[2] - cy[2], l2);
            }
            // W of this node (all three systems) and the velocity stage values
            double W0[5], W1[5], W2[5];
#pragma unroll
            for (int c = 0; c < 4; ++c) { W0[c] = K4L(nb, O_W + c); W1[c] = K4L(nb, O_W + 5 + c); W2[c] = K4L(nb, O_W + 10 + c); }
            const double yv2 = K4L(nb, O_Y + 2), yv3 = K4L(nb, O_Y + 3);
            const double m = node_m(k, i), im = node_im(k, i);
            double fr[5], fcr[5], fci[5];
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                double F[5];
                F[0] = yv2 + (s == 0 ? stage_z<0>(W0[2], W1[2], W2[2]) : s == 1 ? stage_z<1>(W0[2], W1[2], W2[2]) : stage_z<2>(W0[2], W1[2], W2[2]));
                F[1] = yv3 + (s == 0 ? stage_z<0>(W0[3], W1[3], W2[3]) : s == 1 ? stage_z<1>(W0[3], W1[3], W2[3]) : stage_z<2>(W0[3], W1[3], W2[3]));
                if (last) { F[2] = im * cur[s].a; F[3] = im * (cur[s].b - m * GRAV); }
                else { F[2] = im * (cur[s].a - nxt[s].a); F[3] = im * (cur[s].b - nxt[s].b - m * GRAV); }
                F[4] = cur[s].rs - k.r0s;
                const double a0 = (s == 0) ? TI00 : (s == 1) ? TI01 : TI02;
                const double a1 = (s == 0) ? TI10 : (s == 1) ? TI11 : TI12;
                const double a2 = (s == 0) ? TI20 : (s == 1) ? TI21 : TI22;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    finite = finite && is_finite(F[c]);
                    if (s == 0) { fr[c] = a0 * F[c]; fcr[c] = a1 * F[c]; fci[c] = a2 * F[c]; }
                    else { fr[c] = fma(a0, F[c], fr[c]); fcr[c] = fma(a1, F[c], fcr[c]); fci[c] = fma(a2, F[c], fci[c]); }
                }
            }
            // residuals  TI F - mu M W  (mass diag(1,1,1,1,0)); no row scaling in the chain spec
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                fr[c] = fma(-mr, W0[c], fr[c]);
                double br = fma(-mcr, W1[c], fcr[c]); br = fma(mci, W2[c], br);
                double bi = fma(-mcr, W2[c], fci[c]); bi = fma(-mci, W1[c], bi);
                fcr[c] = br; fci[c] = bi;
            }
            fwd_real<STRIDE>(nb, k, i, mrf, fr, zp);
            {   // complex system: g = r_v + mu r_p, g -= L z_prev, z = V g
                double gr[3] = {fma(-mcif, fci[0], fma(mcrf, fcr[0], fcr[2])), fma(-mcif, fci[1], fma(mcrf, fcr[1], fcr[3])), fcr[4]};
                double gi[3] = {fma(mcif, fcr[0], fma(mcrf, fci[0], fci[2])), fma(mcif, fcr[1], fma(mcrf, fci[1], fci[3])), fci[4]};
                if (i > 0) {
                    const double adx = K4L(nb, O_ND + ND_ADX), ady = K4L(nb, O_ND + ND_ADY), bdy = K4L(nb, O_ND + ND_BDY);
                    const double L[3][2] = {{im * adx, im * ady}, {im * ady, im * bdy}, {K4L(nb, O_ND + ND_DX2), K4L(nb, O_ND + ND_DY2)}};
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        gr[q] = fma(-L[q][1], zpr[1], fma(-L[q][0], zpr[0], gr[q]));
                        gi[q] = fma(-L[q][1], zpi[1], fma(-L[q][0], zpi[0], gi[q]));
                    }
                }
                double zr[3], zi[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const double v0r = K4L(nb, O_VCR + 3 * q), v0i = K4L(nb, O_VCI + 3 * q);
                    double sr = fma(v0r, gr[0], -(v0i * gi[0]));
                    double si = fma(v0r, gi[0], v0i * gr[0]);
#pragma unroll
                    for (int c = 1; c < 3; ++c) {
                        const double vr = K4L(nb, O_VCR + 3 * q + c), vi = K4L(nb, O_VCI + 3 * q + c);
                        sr = fma(vr, gr[c], sr); sr = fma(-vi, gi[c], sr);
                        si = fma(vr, gi[c], si); si = fma(vi, gr[c], si);
                    }
                    zr[q] = sr; zi[q] = si;
                    K4W(nb, O_Z + 3 + q) = sr; K4W(nb, O_Z + 6 + q) = si;
                }
                K4W(nb, O_RP + 2) = fcr[0]; K4W(nb, O_RP + 3) = fcr[1]; K4W(nb, O_RP + 4) = fci[0]; K4W(nb, O_RP + 5) = fci[1];
                zpr[0] = zr[0]; zpr[1] = zr[1]; zpi[0] = zi[0]; zpi[1] = zi[1];
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) { cx[s] = nx[s]; cy[s] = ny[s]; cur[s] = nxt[s]; }
        }
    }
    const unsigned mb = warp_ballot_in(m, finite);        // the lanes of this pass whose stage values are finite go on
    if (!finite) return false;
    {   // ---------- backward: block back substitution, velocities, norm, W += dW
        double u[3] = {0, 0, 0}, ur[3] = {0, 0, 0}, ui[3] = {0, 0, 0};
        double s5[5] = {0, 0, 0, 0, 0};
        double rn[5] = {0, 0, 0, 0, 0};
        S.begin(R, O_W, SPN - O_W, k.nn - 1, -1, k.nn, mb);
_Pragma(K4_NEWTON_UNROLL)
        for (int i = k.nn - 1; i >= 0; --i) {
            S.step(i);
            const NodeRef nb = S.ref(i);
            double d0[5], d1[5], d2[5];
            bwd_real<STRIDE>(nb, rn, k, i, mrf, u, d0);
            {
                double nr[3], ni[3];
                if (i == k.nn - 1) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) { nr[q] = K4L(nb, O_Z + 3 + q); ni[q] = K4L(nb, O_Z + 6 + q); }
                } else {
                    const double im = node_im(k, i);
                    const double adx = rn[0], ady = rn[1], bdy = rn[2];
                    const double U[2][3] = {{im * adx, im * ady, im * rn[3]}, {im * ady, im * bdy, im * rn[4]}};
                    const double w0r = fma(U[0][2], ur[2], fma(U[0][1], ur[1], U[0][0] * ur[0]));
                    const double w0i = fma(U[0][2], ui[2], fma(U[0][1], ui[1], U[0][0] * ui[0]));
                    const double w1r = fma(U[1][2], ur[2], fma(U[1][1], ur[1], U[1][0] * ur[0]));
                    const double w1i = fma(U[1][2], ui[2], fma(U[1][1], ui[1], U[1][0] * ui[0]));
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        double xr = K4L(nb, O_Z + 3 + q), xi = K4L(nb, O_Z + 6 + q);
                        const double v0r = K4L(nb, O_VCR + 3 * q), v0i = K4L(nb, O_VCI + 3 * q);
                        const double v1r = K4L(nb, O_VCR + 3 * q + 1), v1i = K4L(nb, O_VCI + 3 * q + 1);
                        xr = fma(-v0r, w0r, xr); xr = fma(v0i, w0i, xr);
                        xi = fma(-v0r, w0i, xi); xi = fma(-v0i, w0r, xi);
                        xr = fma(-v1r, w1r, xr); xr = fma(v1i, w1i, xr);
                        xi = fma(-v1r, w1i, xi); xi = fma(-v1i, w1r, xi);
                        nr[q] = xr; ni[q] = xi;
                    }
                }
                const double r0r = K4L(nb, O_RP + 2), r1r = K4L(nb, O_RP + 3), r0i = K4L(nb, O_RP + 4), r1i = K4L(nb, O_RP + 5);
                d1[0] = nr[0]; d1[1] = nr[1]; d1[4] = nr[2];
                d2[0] = ni[0]; d2[1] = ni[1]; d2[4] = ni[2];
                d1[2] = fma(-mcif, ni[0], fma(mcrf, nr[0], -r0r)); d2[2] = fma(mcif, nr[0], fma(mcrf, ni[0], -r0i));
                d1[3] = fma(-mcif, ni[1], fma(mcrf, nr[1], -r1r)); d2[3] = fma(mcif, nr[1], fma(mcrf, ni[1], -r1i));
#pragma unroll
                for (int q = 0; q < 3; ++q) { ur[q] = nr[q]; ui[q] = ni[q]; }
            }
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const double sc = K4L(nb, O_NS + c);
                double x = d0[c] * sc; s5[c] = fma(x, x, s5[c]);
                x = d1[c] * sc; s5[c] = fma(x, x, s5[c]);
                x = d2[c] * sc; s5[c] = fma(x, x, s5[c]);
                K4W(nb, O_W + c) = K4L(nb, O_W + c) + d0[c]; K4W(nb, O_W + 5 + c) = K4L(nb, O_W + 5 + c) + d1[c]; K4W(nb, O_W + 10 + c) = K4L(nb, O_W + 10 + c) + d2[c];
            }
            rn[0] = K4L(nb, O_ND + ND_ADX); rn[1] = K4L(nb, O_ND + ND_ADY); rn[2] = K4L(nb, O_ND + ND_BDY); rn[3] = K4L(nb, O_ND + ND_AL); rn[4] = K4L(nb, O_ND + ND_BL);       // for the visit of node i - 1
        }
        ss = ((s5[0] + s5[1]) + (s5[2] + s5[3])) + s5[4];
    }
    return true;
}

// attempt set-up in ONE sweep: the Newton scales (:564-588, when the step size is new) and the starting guess W = TI Z0
// (:580-583), Z0 from the previous step's cubic evaluated beyond its interval
template <int STRIDE, class STG>
BRDAE_HD void setup_sweep(STG &S, const Rope<STRIDE> &R, const Consts &k, const SolveArgs &a, bool scale, bool zero, double t, double h,
                          double t_sol_old, double inv_h_sol, unsigned m) {
    double x[3], x2[3], x3[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const double tt = t + h * stage_c(s);
        x[s] = (tt - t_sol_old) * inv_h_sol; x2[s] = x[s] * x[s]; x3[s] = x2[s] * x[s];
    }
    S.begin(R, O_Q, O_W - O_Q, 0, 1, k.nn, m);
    for (int i = 0; i < k.nn; ++i) {
        S.step(i);
        const NodeRef nb = S.ref(i);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const double yc = K4L(nb, O_Y + c);
            if (scale) {
                const int vi = a.var_index_dev[5 * i + c];
                const int ve = vi > 1 ? vi - 1 : 0;
                const double hp = a.scale_newton_norm ? hpow(h, ve) : 1.0;
                K4W(nb, O_NS + c) = hp / (a.atol_dev[5 * i + c] + fabs(yc) * a.rtol);
            }
            if (zero) { K4W(nb, O_W + c) = 0.0; K4W(nb, O_W + 5 + c) = 0.0; K4W(nb, O_W + 10 + c) = 0.0; continue; }
            const double q0 = K4L(nb, O_Q + 3 * c), q1 = K4L(nb, O_Q + 3 * c + 1), q2 = K4L(nb, O_Q + 3 * c + 2);
            const double yo = K4L(nb, O_YO + c);
            double z[3];
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                double v = q0 * x[s];
                v = fma(q1, x2[s], v);
                v = fma(q2, x3[s], v);
                v = v + yo;
                z[s] = v - yc;
            }
            K4W(nb, O_W + c) = fma(TI02, z[2], fma(TI01, z[1], TI00 * z[0]));
            K4W(nb, O_W + 5 + c) = fma(TI12, z[2], fma(TI11, z[1], TI10 * z[0]));
            K4W(nb, O_W + 10 + c) = fma(TI22, z[2], fma(TI21, z[1], TI20 * z[0]));
        }
    }
}

// error estimate (:669-711): pass 0 solves with f + M ZE, pass 1 with rhs(y + err) + M ZE; returns the sum of squares of
// the scaled error.  The reference applies the factors of the hscale-d matrix to the unscaled vector: the rows
// of algebraic unknowns are multiplied by the h of the factorisation.
// `pass` is the same for every lane of the call (mask m).
template <int STRIDE, class STG>
BRDAE_HD double error_sweeps(STG &S, const Rope<STRIDE> &R, const Consts &k, const SolveArgs &a, int pass, double h, double inv_h,
                             double h_lu, double mrf, unsigned m) {
    double zp[2] = {0, 0};
    auto row_scale = [&](int i, double (&e)[5]) {
#pragma unroll
        for (int c = 0; c < 5; ++c)
            if (a.scale_residuals && a.var_index_dev[5 * i + c] >= 1) e[c] = e[c] * h_lu;
    };
    if (pass == 0) {
        S.begin(R, O_F, O_VCR - O_F, 0, 1, k.nn, m);
        for (int i = 0; i < k.nn; ++i) {
            S.step(i);
            const NodeRef nb = S.ref(i);
            double e[5];
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const double w0 = K4L(nb, O_W + c), w1 = K4L(nb, O_W + 5 + c), w2 = K4L(nb, O_W + 10 + c);
                double v = stage_z<0>(w0, w1, w2) * RE0;
                v = fma(stage_z<1>(w0, w1, w2), RE1, v);
                v = fma(stage_z<2>(w0, w1, w2), RE2, v);
                const double mz = (c < 4) ? v * inv_h : 0.0;             // M . ZE
                if (c < 4) K4W(nb, O_MZE + c) = mz;
                e[c] = K4L(nb, O_F + c) + mz;
            }
            row_scale(i, e);
            fwd_real<STRIDE>(nb, k, i, mrf, e, zp);
        }
    } else {
        // err <- rhs(t, y + err) + M ZE, walked forward with the rod of the next node as look-ahead (rhs_sweep's walk)
        double x = 0, y = 0;
        Rod cur;
        S.begin(R, O_Y, O_NS - O_Y, 0, 1, k.nn, m);
        S.step(0);
        {
            const NodeRef nb = S.ref(0);
            x = K4L(nb, O_Y) + K4L(nb, O_ERR); y = K4L(nb, O_Y + 1) + K4L(nb, O_ERR + 1);
            cur = rod(x, y, K4L(nb, O_Y + 4) + K4L(nb, O_ERR + 4));
        }
        for (int i = 0; i < k.nn; ++i) {
            S.step(i);
            const NodeRef nb = S.ref(i);
            const bool last = (i == k.nn - 1);
            Rod nxt = cur;
            double xn = 0, yn = 0;
            if (!last) {
                const NodeRef nn_ = S.ref(i + 1);
                xn = K4L(nn_, O_Y) + K4L(nn_, O_ERR); yn = K4L(nn_, O_Y + 1) + K4L(nn_, O_ERR + 1);
                const double ln = K4L(nn_, O_Y + 4) + K4L(nn_, O_ERR + 4);
                nxt = rod(xn - x, yn - y, ln);
            }
            const double vx = K4L(nb, O_Y + 2) + K4L(nb, O_ERR + 2), vy = K4L(nb, O_Y + 3) + K4L(nb, O_ERR + 3);
            const double m = node_m(k, i), im = node_im(k, i);
            double e[5];
            e[0] = vx; e[1] = vy;
            if (last) { e[2] = im * cur.a; e[3] = im * (cur.b - m * GRAV); }
            else { e[2] = im * (cur.a - nxt.a); e[3] = im * (cur.b - nxt.b - m * GRAV); }
            e[4] = cur.rs - k.r0s;
#pragma unroll
            for (int c = 0; c < 4; ++c) e[c] = e[c] + K4L(nb, O_MZE + c);
            e[4] = e[4] + 0.0;
            row_scale(i, e);
            fwd_real<STRIDE>(nb, k, i, mrf, e, zp);
            x = xn; y = yn; cur = nxt;
        }
    }
    double u[3] = {0, 0, 0}, s5[5] = {0, 0, 0, 0, 0}, rn[5] = {0, 0, 0, 0, 0};
    S.begin(R, O_Y, O_RP + 2 - O_Y, k.nn - 1, -1, k.nn, m);
    for (int i = k.nn - 1; i >= 0; --i) {
        S.step(i);
        const NodeRef nb = S.ref(i);
        double d[5];
        bwd_real<STRIDE>(nb, rn, k, i, mrf, u, d);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int vi = a.var_index_dev[5 * i + c];
            double sq;
            if (a.zero_algebraic_error && vi != 0) { d[c] = 0.0; sq = 0.0; }
            else {
                const double yc = K4L(nb, O_Y + c);
                const double ynew = yc + (K4L(nb, O_W + c) + K4L(nb, O_W + 5 + c));
                const int ve = vi > 1 ? vi - 1 : 0;
                const double hp = a.scale_error ? hpow(h, ve) : 1.0;
                sq = d[c] * (hp / (a.atol_dev[5 * i + c] + fmax(fabs(yc), fabs(ynew)) * a.rtol));
            }
            s5[c] = fma(sq, sq, s5[c]);
            K4W(nb, O_ERR + c) = d[c];
        }
        rn[0] = K4L(nb, O_ND + ND_ADX); rn[1] = K4L(nb, O_ND + ND_ADY); rn[2] = K4L(nb, O_ND + ND_BDY); rn[3] = K4L(nb, O_ND + ND_AL); rn[4] = K4L(nb, O_ND + ND_BL);
    }
    return ((s5[0] + s5[1]) + (s5[2] + s5[3])) + s5[4];
}

// accepted step (:742-784): y_old = y, y += Z_2, Q = Z^T P, f = rhs(y), optionally the Jacobian's rod quantities at the new y
// `out` != nullptr: the first output point of this step is evaluated here, from the coefficients while they are in registers
// (out[(5 i + c) out_stride] = y_old + Q [x, x^2, x^3], the arithmetic of the output loop of the lane loop)
template <int STRIDE, class STG>
BRDAE_HD void accept_sweep(STG &S, const Rope<STRIDE> &R, const Consts &k, bool recompute, unsigned m, double *out, int64_t out_stride,
                           double ox, double ox2, double ox3) {
    // new state of a node: y + (W0 + W1)
    auto ynew = [&](const NodeRef &nb, int c) { return K4L(nb, O_Y + c) + (K4L(nb, O_W + c) + K4L(nb, O_W + 5 + c)); };
    double x, y, l, dx, dy;
    Rod cur;
    S.begin(R, O_Y, O_ND - O_Y, 0, 1, k.nn, m);
    S.step(0);
    {
        const NodeRef nb = S.ref(0);
        x = ynew(nb, 0); y = ynew(nb, 1); l = ynew(nb, 4);
        dx = x; dy = y;
        cur = rod(dx, dy, l);
    }
    for (int i = 0; i < k.nn; ++i) {
        S.step(i);
        const NodeRef nb = S.ref(i);
        const bool last = (i == k.nn - 1);
        Rod nxt = cur;
        double xn = 0, yn = 0, ln = 0, dxn = 0, dyn = 0;
        if (!last) {
            const NodeRef nn_ = S.ref(i + 1);
            xn = ynew(nn_, 0); yn = ynew(nn_, 1); ln = ynew(nn_, 4);
            dxn = xn - x; dyn = yn - y;
            nxt = rod(dxn, dyn, ln);
        }
        double yn5[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const double w0 = K4L(nb, O_W + c), w1 = K4L(nb, O_W + 5 + c), w2 = K4L(nb, O_W + 10 + c);
            const double z0 = stage_z<0>(w0, w1, w2), z1 = stage_z<1>(w0, w1, w2), z2 = stage_z<2>(w0, w1, w2);
            const double yo = K4L(nb, O_Y + c);
            K4W(nb, O_YO + c) = yo;
            yn5[c] = yo + z2;
            K4W(nb, O_Y + c) = yn5[c];
            const double q0 = fma(z2, P20, fma(z1, P10, z0 * P00));
            const double q1 = fma(z2, P21, fma(z1, P11, z0 * P01));
            const double q2 = fma(z2, P22, fma(z1, P12, z0 * P02));
            K4W(nb, O_Q + 3 * c) = q0;
            K4W(nb, O_Q + 3 * c + 1) = q1;
            K4W(nb, O_Q + 3 * c + 2) = q2;
            if (out) {
                double v = q0 * ox;
                v = fma(q1, ox2, v);
                v = fma(q2, ox3, v);
                out[(int64_t)(5 * i + c) * out_stride] = v + yo;
            }
        }
        const double m = node_m(k, i), im = node_im(k, i);
        K4W(nb, O_F) = yn5[2];
        K4W(nb, O_F + 1) = yn5[3];
        if (last) { K4W(nb, O_F + 2) = im * cur.a; K4W(nb, O_F + 3) = im * (cur.b - m * GRAV); }
        else { K4W(nb, O_F + 2) = im * (cur.a - nxt.a); K4W(nb, O_F + 3) = im * (cur.b - nxt.b - m * GRAV); }
        K4W(nb, O_F + 4) = cur.rs - k.r0s;
        if (recompute) store_node_jac<STRIDE>(nb.st, dx, dy, l);
        x = xn; y = yn; l = ln; dx = dxn; dy = dyn; cur = nxt;
    }
}

// ---- the per-lane integrator ----
// `S` stages the node stream of the sweeps (Stager); every sweep is entered by exactly the lanes of a ballot taken by the whole
// warp right before it, and the blocks are separated by full-warp convergence points, so no two sweeps of a warp ever overlap
// (they share the warp's ring of node buffers).
template <int STRIDE, class STG>
BRDAE_HD void k4_lane_loop(const SolveArgs &a, const Rope<STRIDE> &R, bool have_slot, STG &S) {
    const int N = a.n;
    Consts kc;
    kc.nn = N / 5; kc.m_in = 1.0 / (kc.nn - 1); kc.im_in = 1 / kc.m_in;
    { const double r0 = 2.0 / kc.nn; kc.r0s = r0 * r0; }
    const int nn = kc.nn;
    const int MAXIT = a.maxit;
    const double dir = a.dir;

    int state = ST_IDLE, setup_kind = SETUP_STEP;
    long long idx = -1;
    double t = 0, h_abs_state = 0, h_abs_old_state = 0, err_old_state = 0;
    bool hist_state = false, current_jac = true, lu_valid = false, have_sol = false;
    double inv_h_sol = 1, t_sol_old = 0, h_lu = 1.0, mrf = 0, mcrf = 0, mcif = 0;
    int cnt[CN_COUNT];
    int ei = 0;
    long long nsteps = 0;
    double min_step = 0, h_abs = 0, h_abs_old = 0, err_old = 0;
    bool hist = false, rejected = false;
    double h = 0, inv_h = 0, t_new = 0, mr = 0, mcr = 0, mci = 0;
    int kit = 0, nbad = 0, n_iter = 0;
    double dwn_old = 0, rate = 0, safety = 0;
    bool have_old = false, have_rate = false;
#pragma unroll
    for (int q = 0; q < CN_COUNT; ++q) cnt[q] = 0;

    for (;;) {
        int finish_status = RUNNING;
        // ---------------- refill
        if (warp_any(state == ST_IDLE)) {
            const long long got = fetch_system(state == ST_IDLE && have_slot, a.work_counter);
            if (state == ST_IDLE) {
                if (!have_slot || got >= a.B) state = ST_EXIT;
                else {
                    idx = got;
                    t = a.t0;
                    for (int i = 0; i < nn; ++i) {
                        double *nb = R.node(i);
#pragma unroll
                        for (int c = 0; c < 5; ++c) K4G(nb, O_Y + c) = a.y0[(int64_t)(5 * i + c) * a.B + idx];
                    }
                    rhs_sweep(kc, [&](int i, int c) { const double *nb = R.node(i); return K4G(nb, O_Y + c); },
                              [&](int i, int c, double v) { double *nb = R.node(i); K4G(nb, O_F + c) = v; },
                              [&](int i, double dx, double dy, double l) { store_node_jac<STRIDE>(R.node(i), dx, dy, l); });
#pragma unroll
                    for (int q = 0; q < CN_COUNT; ++q) cnt[q] = 0;
                    cnt[CN_NFEV] = 1; cnt[CN_NJEV] = 1;
                    h_abs_state = a.h0;
                    hist_state = false; current_jac = true; lu_valid = false; have_sol = false;
                    h_lu = 1.0; ei = 0; nsteps = 0;
                    if (a.rep_cap > 0) a.rep_count[idx] = 0;
                    state = ST_SETUP; setup_kind = SETUP_STEP;
                    if (t == a.tb) {
                        for (; ei < a.T; ++ei)
                            for (int i = 0; i < nn; ++i) {
                                const double *nb = R.node(i);
#pragma unroll
                                for (int c = 0; c < 5; ++c) a.y_eval[((int64_t)ei * N + 5 * i + c) * a.B + idx] = K4G(nb, O_Y + c);
                            }
                        finish_status = BRDAE_STATUS_FINISHED;
                        state = ST_IDLE;
                    }
                }
            }
        }
        warp_converge();
        if (!warp_any(state != ST_EXIT)) break;

        // ---------------- error estimate, accept/reject, epilogue (:660-784)
        if (warp_any(state == ST_ERREST)) {
            const bool in_e = state == ST_ERREST;
            double error_norm = 0.0, err1 = NAN, err2 = NAN;
            // first estimate: every lane of the block; second estimate (:690-711): the lanes that ask for it
            const bool est1 = in_e && !a.constant_dt;
            const unsigned m1 = warp_ballot(est1);
            if (est1) {
                const double ssq = error_sweeps<STRIDE>(S, R, kc, a, 0, h, inv_h, h_lu, mrf, m1);
                cnt[CN_NSOLVE_ERR]++;
                error_norm = err1 = sqrt(ssq) * (a.zero_algebraic_error ? a.rsq_nondiff : a.rsqn);
            }
            warp_converge();
            const bool est2 = est1 && error_norm > 1 && (a.always_2nd || rejected);
            const unsigned m2 = warp_ballot(est2);
            if (est2) {
                const double ssq = error_sweeps<STRIDE>(S, R, kc, a, 1, h, inv_h, h_lu, mrf, m2);
                cnt[CN_NSOLVE_ERR]++;
                cnt[CN_NFEV]++;
                error_norm = err2 = sqrt(ssq) * (a.zero_algebraic_error ? a.rsq_nondiff : a.rsqn);
            }
            warp_converge();
            bool accepted = in_e, recompute = false;
            double factor = 1;
            if (est1) {
                if (a.rep_cap > 0) report_event(a, idx, error_norm > 1 ? RC_REJECTED : RC_ACCEPTED, t, h, n_iter, nbad, err1, err2);
                if (error_norm > 1) {
                    const double fr = gustafsson(h_abs, h_abs_old, error_norm, err_old, hist, a.predictive_controller);
                    h_abs *= fmax(a.min_factor, safety * fr);
                    lu_valid = false; rejected = true; cnt[CN_NREJ]++;
                    accepted = false;
                    state = ST_SETUP; setup_kind = SETUP_ATTEMPT;
                } else cnt[CN_NACC]++;
            }
            if (accepted) {
                recompute = n_iter > 2 && have_rate && rate > a.jac_recompute;
                if (a.constant_dt) factor = a.max_step / h_abs;
                else factor = fmin(a.max_factor, safety * gustafsson(h_abs, h_abs_old, error_norm, err_old, hist, a.predictive_controller));
                if (!recompute && factor < 1.2) factor = 1; else lu_valid = false;
            }
            const unsigned ma = warp_ballot(accepted);
            if (accepted) {
                // an output point inside this step (usually at most one) is produced by the sweep itself; reading the
                // coefficients back from global memory for it was 15 % of the stall samples (profiles/r02_k4_ncu_summary.md)
                double *out = nullptr;
                double ox = 0, ox2 = 0, ox3 = 0;
                if (ei < a.T) {
                    const double te = a.t_eval[ei];
                    if (dir > 0 ? te <= t_new : te >= t_new) {
                        out = a.y_eval + (int64_t)ei * N * a.B + idx;
                        ox = (te - t) * inv_h; ox2 = ox * ox; ox3 = ox2 * ox;
                    }
                }
                accept_sweep<STRIDE>(S, R, kc, recompute, ma, out, a.B, ox, ox2, ox3);
                if (out) ++ei;
                cnt[CN_NFEV]++;
                if (recompute) { cnt[CN_NJEV]++; current_jac = true; }
                else current_jac = false;
                h_abs_old_state = h_abs_state; err_old_state = error_norm; hist_state = true;
                h_abs_state = h_abs * factor;
                t_sol_old = t; inv_h_sol = inv_h; have_sol = true;
                t = t_new;
                if (a.dense_cap > 0 && nsteps <= a.dense_cap) {      // OdeSolution record of this step (brdae.h)
                    const int64_t ks = nsteps - 1;
                    a.dense_t[ks * a.B + idx] = t;
                    for (int i = 0; i < nn; ++i) {
                        const double *nb = R.node(i);
#pragma unroll
                        for (int c = 0; c < 5; ++c) {
                            a.dense_y[(ks * N + 5 * i + c) * a.B + idx] = K4G(nb, O_Y + c);
#pragma unroll
                            for (int j = 0; j < 3; ++j) a.dense_q[((ks * N + 5 * i + c) * 3 + j) * a.B + idx] = K4G(nb, O_Q + 3 * c + j);
                        }
                    }
                }
                while (ei < a.T) {
                    const double te = a.t_eval[ei];
                    if (!(dir > 0 ? te <= t : te >= t)) break;
                    const double x = (te - t_sol_old) * inv_h_sol;
                    const double x2 = x * x, x3 = x2 * x;
                    for (int i = 0; i < nn; ++i) {
                        const double *nb = R.node(i);
#pragma unroll
                        for (int c = 0; c < 5; ++c) {
                            double v = K4G(nb, O_Q + 3 * c) * x;
                            v = fma(K4G(nb, O_Q + 3 * c + 1), x2, v);
                            v = fma(K4G(nb, O_Q + 3 * c + 2), x3, v);
                            a.y_eval[((int64_t)ei * N + 5 * i + c) * a.B + idx] = v + K4G(nb, O_YO + c);
                        }
                    }
                    ++ei;
                }
                if (dir * (t - a.tb) >= 0) finish_status = BRDAE_STATUS_FINISHED;
                else { state = ST_SETUP; setup_kind = SETUP_STEP; }
            }
        }
        warp_converge();
        // ---------------- cheap set-up: prologue :516-562, attempt :564-588, guess :580-583
        if (warp_any(state == ST_SETUP && finish_status == RUNNING)) {
            bool new_scales = false;
            if (state == ST_SETUP && finish_status == RUNNING) {
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
                if (finish_status == RUNNING && setup_kind <= SETUP_ATTEMPT) {
                    if (h_abs < min_step) finish_status = BRDAE_STATUS_TOO_SMALL_STEP;
                    else {
                        h = h_abs * dir;
                        t_new = t + h;
                        if (dir * (t_new - a.tb) > 0) t_new = a.tb;
                        h = t_new - t;
                        h_abs = fabs(h);
                        inv_h = 1.0 / h;
                        new_scales = true;
                        mr = MU_REAL * inv_h; mcr = MU_CR * inv_h; mci = MU_CI * inv_h;
                    }
                }
            }
            const bool go = state == ST_SETUP && finish_status == RUNNING;
            const unsigned ms = warp_ballot(go);
            if (go) {
                setup_sweep<STRIDE>(S, R, kc, a, new_scales, !have_sol || !a.extrapolate, t, h, t_sol_old, inv_h_sol, ms);
                kit = 0; nbad = 0; have_old = false; have_rate = false; rate = 0; dwn_old = 0;
                state = lu_valid ? ST_NEWTON : ST_FACTOR;
            }
        }
        warp_converge();
        // ---------------- retire
        if (warp_any(finish_status != RUNNING)) {
            if (finish_status != RUNNING) {
                for (int e = ei; e < a.T; ++e)
                    for (int c = 0; c < N; ++c) a.y_eval[((int64_t)e * N + c) * a.B + idx] = NAN;
                for (int i = 0; i < nn; ++i) {
                    const double *nb = R.node(i);
#pragma unroll
                    for (int c = 0; c < 5; ++c) a.y_final[(int64_t)(5 * i + c) * a.B + idx] = K4G(nb, O_Y + c);
                }
                a.n_eval[idx] = ei; a.t_final[idx] = t; a.status[idx] = finish_status;
#pragma unroll
                for (int q = 0; q < CN_COUNT; ++q) a.counters[(int64_t)q * a.B + idx] = cnt[q];
                state = ST_IDLE;
            }
        }
        // ---------------- (re)factorisation :594-613
        warp_converge();
        if (warp_any(state == ST_FACTOR)) {
            const unsigned mf = warp_ballot(state == ST_FACTOR);
            if (state == ST_FACTOR) {
                if (a.scale_residuals) h_lu = h;
                mrf = mr; mcrf = mcr; mcif = mci;
                factor_sweep<STRIDE>(S, R, kc, mr, mcr, mci, mf);
                cnt[CN_NLU]++;
                if (a.rep_cap > 0) report_event(a, idx, RC_FACTORISATION, t, h, -1, -1, -1.0, -1.0);
                lu_valid = true;
                state = ST_NEWTON;
            }
        }
        warp_converge();
        // ---------------- simplified-Newton iterations :846-940
        // At least a.newton_reps passes (while any rope iterates); beyond that for as long as more than newton_vote / 32 of
        // the live ropes of the warp still iterate: the chain refactorises at every step and needs 4-5 iterations per
        // attempt, so with a fixed count most ropes wait a whole trip of the other blocks for their last iteration
        // (profiles/tools/k4_sched_sim.py: lane efficiency 0.42 -> 0.56).  Scheduling only: results do not depend on it.
#pragma unroll 1
        for (int rep = 0;; ++rep) {
            const int iterating = warp_count(state == ST_NEWTON);
            if (iterating == 0) break;
            if (rep >= a.newton_reps && (a.newton_vote <= 0 || 32 * iterating <= a.newton_vote * warp_count(state != ST_EXIT))) break;
            bool need_jac = false;
            const unsigned mn = warp_ballot(state == ST_NEWTON);
            {
                if (state == ST_NEWTON) {
                    double ssq = 0;
                    const bool finite = newton_sweeps<STRIDE>(S, R, kc, mr, mcr, mci, mrf, mcrf, mcif, ssq, mn);
                    cnt[CN_NFEV] += 3;
                    int outcome = 0;
                    bool update_old = true;
                    if (!finite) outcome = 2;
                    else {
                        cnt[CN_NSOLVE]++;
                        const double dwn = sqrt(ssq) * a.rsq3n;
                        double dw_true = 0, dw_true_max = 0;
                        if (have_old) {
                            rate = dwn / dwn_old; have_rate = true;
                            const double inv1 = 1.0 / (1.0 - rate);
                            dw_true = rate * inv1 * dwn;
                            double rp = rate;
                            for (int e = 1; e < MAXIT; ++e) rp = (e < MAXIT - kit) ? rp * rate : rp;
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
                        ++kit;
                        if (kit >= MAXIT) { outcome = 2; n_iter = MAXIT; }
                    } else n_iter = kit + 1;
                    if (outcome != 0) {
                        safety = a.safety_factor * (2 * MAXIT + 1) / (2 * MAXIT + n_iter);
                        if (outcome == 1) state = ST_ERREST;
                        else if (current_jac) {
                            if (a.rep_cap > 0) report_event(a, idx, RC_NON_CONV, t, h, n_iter, nbad, -1.0, -1.0);
                            cnt[CN_NFAIL]++;
                            h_abs *= a.nonconv_factor;
                            lu_valid = false;
                            state = ST_SETUP; setup_kind = SETUP_ATTEMPT;
                        } else need_jac = true;
                    }
                }
            }
            warp_converge();
            if (warp_any(need_jac)) {           // stale Jacobian: re-evaluate it and repeat the attempt (:641-648)
                const unsigned mj = warp_ballot(need_jac);
                if (need_jac) {
                    jac_sweep<STRIDE>(S, R, kc, mj);
                    cnt[CN_NJEV]++;
                    if (a.rep_cap > 0) report_event(a, idx, RC_JACOBIAN_UPDATE, t, h, -1, -1, -1.0, -1.0);
                    current_jac = true; lu_valid = false;
                    state = ST_SETUP; setup_kind = SETUP_GUESS;
                }
                warp_converge();
            }
        }
    }
}

}  // namespace k4
}  // namespace brdae
