// radau_k1.cuh -- K1: thread-per-system batched RadauDAE for small systems (N <= 8), sm_100a.
//
// One lane integrates one system at a time; a persistent grid pulls system indices from a
// global work counter (warp-aggregated atomicAdd: ballot/popc/shfl), so lanes whose trajectory
// is finished, failed or short are refilled at once instead of idling until the slowest lane
// of the warp is done.
//
// The reference's nested loops (radauDAE.py:564 `while not step_accepted`, :591 `while not
// converged`, :846 Newton `for`) are flattened into a per-lane state machine with three
// expensive states -- FACTOR (iteration matrices + both LUs), NEWTON (one simplified-Newton
// iteration) and ERREST (error estimate, accept/reject, controller, dense output) -- and one
// cheap one (SETUP: step prologue, attempt set-up, starting guess).  The blocks are laid out in the
// order a system walks through them and every trip of the loop is one fixed CYCLE
//     ERREST -> SETUP -> retire -> FACTOR -> NEWTON x newton_reps,
// each block executed by exactly the lanes that wait for it (a warp none of whose lanes wants a block
// skips it).  A lane that converged does its error estimate, the next step's set-up, the
// refactorisation if one is due and its next Newton iterations in a single cycle, and because an
// attempt needs 2-4 Newton iterations the lanes of a CTA phase-lock onto the cycle by themselves: no
// vote, one barrier per cycle.  History (pendulum ensemble, 1M systems): every block every trip 301 ms;
// majority vote per block 109 ms; majority entry into the chain 96 ms; + two Newton passes 83 ms;
// fixed cycle with four passes 76 ms.  profiles/tools/sched_sim.py replays the oracle's event traces
// under such policies and predicted each of these steps (lane efficiency 0.56 -> 0.80).
//
// Storage (N = 5): the real and complex LU factors (3 N^2 doubles), f(t,y), y_old, the dense
// output coefficients Q (3N) and the point (t_J, y_J) the Jacobian was last evaluated at live in
// shared memory, lane-strided ([slot][lane]: conflict-free 8-byte accesses); y, W and the
// controller scalars live in registers.  The Jacobian itself is never stored: it is
// re-materialised from (t_J, y_J) when a factorisation needs it, which reproduces the
// reference's "stale J, new h" refactorisations (radauDAE.py:594-613 with self.J kept) exactly.
// The LUs are computed left-looking, one column at a time in registers against the finished
// columns in shared memory, so a factorisation needs N (2N complex) live values instead of N^2.
//
// Reference map: prologue :516-562, attempt :564-588, factorisation :594-613, Newton :793-949,
// non-convergence :635-658, error estimate :669-711, accept/reject :713-739, epilogue :742-784,
// dense output :951-982, driver / t_eval slicing :1079-1157 (scipy ivp.py:659-732).
#pragma once
#include <type_traits>

#include "problems.cuh"

#ifndef BRDAE_K1_IDENT_FASTPATH
#define BRDAE_K1_IDENT_FASTPATH 1    // substitutions on compile-time addresses when no lane of the warp pivoted (0: always the general code)
#endif
#ifndef BRDAE_K1_FIRST_ITER_FASTPATH
#define BRDAE_K1_FIRST_ITER_FASTPATH 1    // skip the rate arithmetic of the Newton convergence test when no lane of the warp has a previous norm
#endif
#ifndef BRDAE_K1_STATIC_LU
#define BRDAE_K1_STATIC_LU 1    // optimistic factorisation on compile-time row addresses for systems that have never pivoted
#endif
#ifndef BRDAE_SIM_COUNT
#define BRDAE_SIM_COUNT(i)      // tests/host_sim counts how often the fast paths and their fall-backs run; nothing on the device
#endif
#ifndef BRDAE_SIM_NEIGHBOUR
#define BRDAE_SIM_NEIGHBOUR() true   // tests/host_sim (a warp of ONE lane) lets an imaginary neighbouring lane veto the warp-wide fast paths at random
#endif
#ifndef K1_TEVAL_SHARED
#define K1_TEVAL_SHARED 256      // output grids up to this many points are staged in shared memory
#endif

namespace brdae {

enum : int { ST_IDLE = 0, ST_SETUP, ST_FACTOR, ST_NEWTON, ST_ERREST, ST_JACFD, ST_EXIT };
// what a lane entering SETUP still has to do
enum : int { SETUP_STEP = 0, SETUP_ATTEMPT = 1, SETUP_GUESS = 2 };
constexpr int RUNNING = 0x7fff;     // finish_status of a lane whose system goes on (not a BRDAE_STATUS_* value)

// lane-strided shared-memory slots of one lane
template <int STRIDE>
struct Slots {
    double *base;
    BRDAE_HD double &operator()(int e) const { return base[e * STRIDE]; }
};

// ---- LU of the spec, left-looking, one column at a time ------------------------------------
// Same arithmetic as a right-looking elimination with partial pivoting on |a| (first maximum) and
// reciprocal pivots: element (row, k) receives its updates -l(row,j) u(j,k), j = 0,1,.. in the same
// order.  Rows are NEVER moved: every row of the factors stays where the matrix row it came from was
// stored, and the row permutation is a packed 3-bit table (position i of the pivoted order -> original
// row).  A pivoted access is then a shared-memory load at `row_of(perm, i)`, a dynamic ADDRESS, which
// is free, while a dynamic REGISTER index would cost a chain of selects per element (the first
// version spent more instructions on selects than on FP64 math) and physically exchanging the rows
// (the second version) cost 120 shared-memory accesses per factorisation pair: the kernel is bound by
// shared-memory bandwidth (2 wavefronts per 8-byte warp access), not by the FP64 pipe.
// `column(k, col)` fills col[r] = A[r][k] in original row order; it is parked in the still-unused
// column k of the factor storage and read back in pivoted order.
BRDAE_HD unsigned perm_identity(int n) {
    unsigned p = 0;
    for (int i = 0; i < n; ++i) p |= (unsigned)i << (3 * i);
    return p;
}
BRDAE_HD unsigned perm_swap(unsigned perm, int k, int p) {
    const unsigned fk = (perm >> (3 * k)) & 7u, fp = (perm >> (3 * p)) & 7u;
    perm &= ~((7u << (3 * k)) | (7u << (3 * p)));
    return perm | (fp << (3 * k)) | (fk << (3 * p));     // k == p: unchanged
}
template <int N>
BRDAE_HD int row_of(unsigned perm, int i) { return (int)((perm >> (3 * i)) & 7u) * N; }   // slot offset of pivoted row i

template <int N, class S, class ColFn>
BRDAE_HD unsigned lu_real_left(const S &sm, int base, ColFn &&column) {
    unsigned perm = perm_identity(N);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double cur[N];
        column(k, cur);
        if (k > 0) {
#pragma unroll
            for (int r = 0; r < N; ++r) sm(base + r * N + k) = cur[r];
#pragma unroll
            for (int i = 0; i < N; ++i) cur[i] = sm(base + row_of<N>(perm, i) + k);
        }
#pragma unroll
        for (int j = 0; j < k; ++j) {
            const double cj = cur[j];
#pragma unroll
            for (int r = j + 1; r < N; ++r) cur[r] = fma(-sm(base + row_of<N>(perm, r) + j), cj, cur[r]);
        }
        int p = k;
        double best = fabs(cur[k]), pv = cur[k];
#pragma unroll
        for (int r = k + 1; r < N; ++r) {
            const double v = fabs(cur[r]);
            const bool g = v > best;
            best = g ? v : best; p = g ? r : p; pv = g ? cur[r] : pv;
        }
        const double rcp = 1.0 / pv;
#pragma unroll
        for (int i = 0; i < N; ++i) sm(base + row_of<N>(perm, i) + k) = (i < k) ? cur[i] : cur[i] * rcp;
        perm = perm_swap(perm, k, p);                       // the multipliers stay with their rows
        sm(base + row_of<N>(perm, k) + k) = rcp;
    }
    return perm;
}
// The same factorisation for a matrix that needs no row exchange -- every pivot candidate search ends on the diagonal,
// which is the rule for the pendulum and Robertson iteration matrices: rows are addressed by compile-time offsets and a
// column never takes the round trip through its parking slots.  Returns false (factors unusable) as soon as a pivot is
// found off the diagonal; the caller then runs the general routine.  When it returns true it has executed exactly the
// arithmetic of the general routine with the identity permutation.
template <int N, class S, class ColFn>
BRDAE_HD bool lu_real_static(const S &sm, int base, ColFn &&column) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double cur[N];
        column(k, cur);
#pragma unroll
        for (int j = 0; j < k; ++j) {
            const double cj = cur[j];
#pragma unroll
            for (int r = j + 1; r < N; ++r) cur[r] = fma(-sm(base + r * N + j), cj, cur[r]);
        }
        const double best = fabs(cur[k]);
#pragma unroll
        for (int r = k + 1; r < N; ++r) ok = ok && !(fabs(cur[r]) > best);
        const double rcp = 1.0 / cur[k];
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i != k) sm(base + i * N + k) = (i < k) ? cur[i] : cur[i] * rcp;
        sm(base + k * N + k) = rcp;
    }
    return ok;
}
template <int N, class S, class ColFn>
BRDAE_HD bool lu_cplx_static(const S &sm, int base_r, int base_i, ColFn &&column) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double cr[N], ci[N];
        column(k, cr, ci);
#pragma unroll
        for (int j = 0; j < k; ++j) {
            const double ur = cr[j], ui = ci[j];
#pragma unroll
            for (int r = j + 1; r < N; ++r) {
                const double lr = sm(base_r + r * N + j), li = sm(base_i + r * N + j);
                double xr = cr[r], xi = ci[r];
                xr = fma(-lr, ur, xr); xr = fma(li, ui, xr);
                xi = fma(-lr, ui, xi); xi = fma(-li, ur, xi);
                cr[r] = xr; ci[r] = xi;
            }
        }
        const double best = fabs(cr[k]) + fabs(ci[k]);
#pragma unroll
        for (int r = k + 1; r < N; ++r) ok = ok && !(fabs(cr[r]) + fabs(ci[r]) > best);
        const double pr = cr[k], pi = ci[k];
        const double invd = 1.0 / fma(pr, pr, pi * pi);
        const double qr = pr * invd, qi = -(pi * invd);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int dst = i * N + k;
            if (i < k) { sm(base_r + dst) = cr[i]; sm(base_i + dst) = ci[i]; }
            else if (i > k) {
                sm(base_r + dst) = fma(cr[i], qr, -(ci[i] * qi));
                sm(base_i + dst) = fma(cr[i], qi, ci[i] * qr);
            }
        }
        sm(base_r + k * N + k) = qr; sm(base_i + k * N + k) = qi;
    }
    return ok;
}
// complex: pivot on |re|+|im| (BLAS izamax), reciprocal pivot conj(p)/|p|^2
template <int N, class S, class ColFn>
BRDAE_HD unsigned lu_cplx_left(const S &sm, int base_r, int base_i, ColFn &&column) {
    unsigned perm = perm_identity(N);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double cr[N], ci[N];
        column(k, cr, ci);
        if (k > 0) {
#pragma unroll
            for (int r = 0; r < N; ++r) { sm(base_r + r * N + k) = cr[r]; sm(base_i + r * N + k) = ci[r]; }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int src = row_of<N>(perm, i) + k;
                cr[i] = sm(base_r + src); ci[i] = sm(base_i + src);
            }
        }
#pragma unroll
        for (int j = 0; j < k; ++j) {
            const double ur = cr[j], ui = ci[j];
#pragma unroll
            for (int r = j + 1; r < N; ++r) {
                const int src = row_of<N>(perm, r) + j;
                const double lr = sm(base_r + src), li = sm(base_i + src);
                double xr = cr[r], xi = ci[r];
                xr = fma(-lr, ur, xr); xr = fma(li, ui, xr);
                xi = fma(-lr, ui, xi); xi = fma(-li, ur, xi);
                cr[r] = xr; ci[r] = xi;
            }
        }
        int p = k;
        double best = fabs(cr[k]) + fabs(ci[k]), pr = cr[k], pi = ci[k];
#pragma unroll
        for (int r = k + 1; r < N; ++r) {
            const double v = fabs(cr[r]) + fabs(ci[r]);
            const bool g = v > best;
            best = g ? v : best; p = g ? r : p; pr = g ? cr[r] : pr; pi = g ? ci[r] : pi;
        }
        const double invd = 1.0 / fma(pr, pr, pi * pi);
        const double qr = pr * invd, qi = -(pi * invd);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int dst = row_of<N>(perm, i) + k;
            if (i < k) { sm(base_r + dst) = cr[i]; sm(base_i + dst) = ci[i]; }
            else {
                sm(base_r + dst) = fma(cr[i], qr, -(ci[i] * qi));
                sm(base_i + dst) = fma(cr[i], qi, ci[i] * qr);
            }
        }
        perm = perm_swap(perm, k, p);
        const int dk = row_of<N>(perm, k) + k;
        sm(base_r + dk) = qr; sm(base_i + dk) = qi;
    }
    return perm;
}

// b <- P b (b[i] = b_old[perm_i]): a round trip through N scratch slots with a dynamic load address, or
// (SELECT, used by the cold-state layout that has no scratch slots) select chains in registers, which trade
// 2N shared-memory accesses for ~N^2 selects at the same speed.
template <int N, bool SELECT, class S>
BRDAE_HD void permute_rhs(const S &sm, int scratch, unsigned perm, double (&b)[N]) {
  if (SELECT) {
    double o[N];
#pragma unroll
    for (int c = 0; c < N; ++c) o[c] = b[c];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const unsigned pi = (perm >> (3 * i)) & 7u;
        double v = o[0];
#pragma unroll
        for (int c = 1; c < N; ++c) v = (pi == (unsigned)c) ? o[c] : v;
        b[i] = v;
    }
  } else {
#pragma unroll
    for (int c = 0; c < N; ++c) sm(scratch + c) = b[c];
#pragma unroll
    for (int i = 0; i < N; ++i) b[i] = sm(scratch + (int)((perm >> (3 * i)) & 7u));
  }
}
// ---- triangular solves streaming the factors from lane-strided shared memory ----------------
// b is put into pivoted order by a round trip through N scratch slots (dynamic load address).
// IDENT: the caller knows that no row was exchanged in this factorisation (perm is the identity for every lane of
// the warp, which is the rule for the pendulum and Robertson families): the right-hand side needs no permutation and
// every factor address is a compile-time offset.  Same loads, same arithmetic, same bits.
template <int N, bool SELECT, bool IDENT = false, class S>
BRDAE_HD void solve_real_sm(const S &sm, int base, int scratch, unsigned perm, double (&b)[N]) {
    if (!IDENT) permute_rhs<N, SELECT>(sm, scratch, perm, b);
    int ro[N];
#pragma unroll
    for (int i = 0; i < N; ++i) ro[i] = base + (IDENT ? i * N : row_of<N>(perm, i));
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double bk = b[k];
#pragma unroll
        for (int r = k + 1; r < N; ++r) b[r] = fma(-sm(ro[r] + k), bk, b[r]);
    }
    // back substitution, column oriented (x_k final once the columns > k have been applied)
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
        const double xk = b[k] * sm(ro[k] + k);
        b[k] = xk;
#pragma unroll
        for (int r = 0; r < k; ++r) b[r] = fma(-sm(ro[r] + k), xk, b[r]);
    }
}
template <int N, bool SELECT, bool IDENT = false, class S>
BRDAE_HD void solve_cplx_sm(const S &sm, int base_r, int base_i, int scratch, unsigned perm, double (&br)[N], double (&bi)[N]) {
    if (!IDENT) {
        permute_rhs<N, SELECT>(sm, scratch, perm, br);
        permute_rhs<N, SELECT>(sm, scratch, perm, bi);
    }
    int ro[N];
#pragma unroll
    for (int i = 0; i < N; ++i) ro[i] = IDENT ? i * N : row_of<N>(perm, i);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double xr = br[k], xi = bi[k];
#pragma unroll
        for (int r = k + 1; r < N; ++r) {
            const double lr = sm(base_r + ro[r] + k), li = sm(base_i + ro[r] + k);
            double yr = br[r], yi = bi[r];
            yr = fma(-lr, xr, yr); yr = fma(li, xi, yr);
            yi = fma(-lr, xi, yi); yi = fma(-li, xr, yi);
            br[r] = yr; bi[r] = yi;
        }
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
        const double qr = sm(base_r + ro[k] + k), qi = sm(base_i + ro[k] + k);
        const double sr = br[k], si = bi[k];
        const double xr = fma(sr, qr, -(si * qi));
        const double xi = fma(sr, qi, si * qr);
        br[k] = xr; bi[k] = xi;
#pragma unroll
        for (int r = 0; r < k; ++r) {
            const double ur = sm(base_r + ro[r] + k), ui = sm(base_i + ro[r] + k);
            double yr = br[r], yi = bi[r];
            yr = fma(-ur, xr, yr); yr = fma(ui, xi, yr);
            yi = fma(-ur, xi, yi); yi = fma(-ui, xr, yi);
            br[r] = yr; bi[r] = yi;
        }
    }
}

// ---- compile-time sparsity of the factors (problems that declare the structure of their Jacobian) ---------------------
// A functor may declare `static BRDAE_CX bool jac_nz(int r, int c)`, the structural non-zeros of its Jacobian.  Together with
// the mass-matrix policy this gives the pattern of mu M - J and, by a symbolic elimination WITHOUT row exchanges, of its LU
// factors (fill included).  The optimistic factorisation and the substitutions on compile-time addresses -- the paths taken
// while nothing pivots -- then skip every operation on a structural zero: multiplying by an exact zero and adding the
// product changes no bit of a finite result, so the factors and solutions are those of the dense routines (the index-3
// pendulum keeps 21 of 45 multiply-adds of a substitution and a third of the elimination).  Entries outside the pattern
// are never written by these paths; they are zeroed when a system is loaded, because the general (pivoting) routines,
// which a neighbouring lane can force on the whole warp, read every entry.
template <class P, class = void> struct trait_jac_nz { static BRDAE_CX bool get(int, int) { return true; } static constexpr bool declared = false; };
template <class P> struct trait_jac_nz<P, decltype((void)P::jac_nz(0, 0))> {
    static BRDAE_CX bool get(int r, int c) { return P::jac_nz(r, c); }
    static constexpr bool declared = true;
};
template <class Prob, class Mass, bool DENSE>
struct LuPattern {
    static constexpr int N = Prob::N;
    bool a[N][N];      // mu M - J
    bool f[N][N];      // L (below the diagonal) and U (on and above it) of the elimination without row exchanges
    BRDAE_CX LuPattern() : a{}, f{} {
        for (int r = 0; r < N; ++r)
            for (int c = 0; c < N; ++c) {
                a[r][c] = DENSE || Mass::nz(r, c) || trait_jac_nz<Prob>::get(r, c);
                f[r][c] = a[r][c] || r == c;
            }
        for (int k = 0; k < N; ++k)
            for (int r = k + 1; r < N; ++r)
                if (f[r][k])
                    for (int c = k + 1; c < N; ++c)
                        if (f[k][c]) f[r][c] = true;
    }
};
template <class Prob, class Mass, bool DENSE>
struct LuPat { static constexpr LuPattern<Prob, Mass, DENSE> v{}; };

template <int I, int E, class F>
BRDAE_HD void static_for(F &&f) {
    if constexpr (I < E) { f(std::integral_constant<int, I>{}); static_for<I + 1, E>(f); }
}
template <int I, int E, class F>
BRDAE_HD void static_for_down(F &&f) {      // I = E-1 ... 0 ... : indices E-1 down to I
    if constexpr (I < E) { f(std::integral_constant<int, E - 1>{}); static_for_down<I, E - 1>(f); }
}

// columns of hscale (mu M - J), real and complex, on the structural non-zeros (entries outside the pattern: exact zeros);
// a functor with explicit recursion rather than nested generic lambdas (those crash cudafe++ 12.9 inside the lane loop)
template <class PAT, class Prob, class Mass, int N>
struct PatColumns {
    const typename Prob::P &prm;
    const SolveArgs &a;
    const double (&J)[N][N];
    const double (&hs)[N];
    double mr, mcr, mci;
    template <int K, int R>
    BRDAE_HD void real_rows(double (&col)[N]) const {
        if constexpr (R < N) {
            if constexpr (!PAT::v.a[R][K]) col[R] = 0.0;
            else {
                const double v = Mass::nz(R, K) ? fma(mr, Mass::get(prm, a, R, K), -J[R][K]) : -J[R][K];
                col[R] = hs[R] * v;
            }
            real_rows<K, R + 1>(col);
        }
    }
    template <int K, int R>
    BRDAE_HD void cplx_rows(double (&cr)[N], double (&ci)[N]) const {
        if constexpr (R < N) {
            if constexpr (!PAT::v.a[R][K]) { cr[R] = 0.0; ci[R] = 0.0; }
            else {
                const bool nz = Mass::nz(R, K);
                const double m = nz ? Mass::get(prm, a, R, K) : 0.0;
                const double v = nz ? fma(mcr, m, -J[R][K]) : -J[R][K];
                cr[R] = hs[R] * v;
                ci[R] = nz ? hs[R] * (mci * m) : 0.0;
            }
            cplx_rows<K, R + 1>(cr, ci);
        }
    }
    template <class KC>
    BRDAE_HD void operator()(KC, double (&col)[N]) const { real_rows<KC::value, 0>(col); }
    template <class KC>
    BRDAE_HD void operator()(KC, double (&cr)[N], double (&ci)[N]) const { cplx_rows<KC::value, 0>(cr, ci); }
};

// zero the factor entries the pattern-aware paths never write (see above); called when a system is loaded
template <class PAT, int N, class S>
BRDAE_HD void zero_outside_pattern(const S &sm, int base_r, int base_cr, int base_ci) {
    static_for<0, N>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        static_for<0, N>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            if constexpr (!PAT::v.f[r][c]) { sm(base_r + r * N + c) = 0.0; sm(base_cr + r * N + c) = 0.0; sm(base_ci + r * N + c) = 0.0; }
        });
    });
}

// lu_real_static / lu_cplx_static restricted to the pattern: `column(kc, cur)` fills the structurally non-zero entries of
// column k (the others are not read).  Same operations, in the same order, on the entries that can be non-zero.
template <class PAT, int N, class S, class ColFn>
BRDAE_HD bool lu_real_static_p(const S &sm, int base, ColFn &&column) {
    bool ok = true;
    static_for<0, N>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        double cur[N];
        column(kc, cur);
        static_for<0, k>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            if constexpr (PAT::v.f[j][k]) {
                const double cj = cur[j];
                static_for<j + 1, N>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    if constexpr (PAT::v.f[r][j]) {
                        cur[r] = fma(-sm(base + r * N + j), cj, cur[r]);
                    }
                });
            }
        });
        const double best = fabs(cur[k]);
        static_for<k + 1, N>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            if constexpr (PAT::v.f[r][k]) ok = ok && !(fabs(cur[r]) > best);
        });
        const double rcp = 1.0 / cur[k];
        static_for<0, N>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (i != k && PAT::v.f[i][k]) sm(base + i * N + k) = (i < k) ? cur[i] : cur[i] * rcp;
        });
        sm(base + k * N + k) = rcp;
    });
    return ok;
}
template <class PAT, int N, class S, class ColFn>
BRDAE_HD bool lu_cplx_static_p(const S &sm, int base_r, int base_i, ColFn &&column) {
    bool ok = true;
    static_for<0, N>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        double cr[N], ci[N];
        column(kc, cr, ci);
        static_for<0, k>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            if constexpr (PAT::v.f[j][k]) {
                const double ur = cr[j], ui = ci[j];
                static_for<j + 1, N>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    if constexpr (PAT::v.f[r][j]) {
                        const double lr = sm(base_r + r * N + j), li = sm(base_i + r * N + j);
                        double xr = cr[r], xi = ci[r];
                        xr = fma(-lr, ur, xr); xr = fma(li, ui, xr);
                        xi = fma(-lr, ui, xi); xi = fma(-li, ur, xi);
                        cr[r] = xr; ci[r] = xi;
                    }
                });
            }
        });
        const double best = fabs(cr[k]) + fabs(ci[k]);
        static_for<k + 1, N>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            if constexpr (PAT::v.f[r][k]) ok = ok && !(fabs(cr[r]) + fabs(ci[r]) > best);
        });
        const double pr = cr[k], pi = ci[k];
        const double invd = 1.0 / fma(pr, pr, pi * pi);
        const double qr = pr * invd, qi = -(pi * invd);
        static_for<0, N>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (PAT::v.f[i][k]) {
                constexpr int dst = i * N + k;
                if constexpr (i < k) { sm(base_r + dst) = cr[i]; sm(base_i + dst) = ci[i]; }
                else if constexpr (i > k) {
                    sm(base_r + dst) = fma(cr[i], qr, -(ci[i] * qi));
                    sm(base_i + dst) = fma(cr[i], qi, ci[i] * qr);
                }
            }
        });
        sm(base_r + k * N + k) = qr; sm(base_i + k * N + k) = qi;
    });
    return ok;
}
// substitutions with identity permutation, restricted to the pattern (solve_real_sm<.., IDENT = true> on the non-zeros)
template <class PAT, int N, class S>
BRDAE_HD void solve_real_ident_p(const S &sm, int base, double (&b)[N]) {
    static_for<0, N>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        const double bk = b[k];
        static_for<k + 1, N>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            if constexpr (PAT::v.f[r][k]) b[r] = fma(-sm(base + r * N + k), bk, b[r]);
        });
    });
    static_for_down<0, N>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        const double xk = b[k] * sm(base + k * N + k);
        b[k] = xk;
        static_for<0, k>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            if constexpr (PAT::v.f[r][k]) b[r] = fma(-sm(base + r * N + k), xk, b[r]);
        });
    });
}
template <class PAT, int N, class S>
BRDAE_HD void solve_cplx_ident_p(const S &sm, int base_r, int base_i, double (&br)[N], double (&bi)[N]) {
    static_for<0, N>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        const double xr = br[k], xi = bi[k];
        static_for<k + 1, N>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            if constexpr (PAT::v.f[r][k]) {
                const double lr = sm(base_r + r * N + k), li = sm(base_i + r * N + k);
                double yr = br[r], yi = bi[r];
                yr = fma(-lr, xr, yr); yr = fma(li, xi, yr);
                yi = fma(-lr, xi, yi); yi = fma(-li, xr, yi);
                br[r] = yr; bi[r] = yi;
            }
        });
    });
    static_for_down<0, N>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        const double qr = sm(base_r + k * N + k), qi = sm(base_i + k * N + k);
        const double sr = br[k], si = bi[k];
        const double xr = fma(sr, qr, -(si * qi));
        const double xi = fma(sr, qi, si * qr);
        br[k] = xr; bi[k] = xi;
        static_for<0, k>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            if constexpr (PAT::v.f[r][k]) {
                const double ur = sm(base_r + r * N + k), ui = sm(base_i + r * N + k);
                double yr = br[r], yi = bi[r];
                yr = fma(-ur, xr, yr); yr = fma(ui, xi, yr);
                yi = fma(-ur, xi, yi); yi = fma(-ui, xr, yi);
                br[r] = yr; bi[r] = yi;
            }
        });
    });
}

// Slots of one lane.  HOT: the LU factors (+ N scratch slots for right-hand sides on their way into
// pivoted order), always in shared memory.  COLD: f(t,y), y_old, the dense-output coefficients Q and the
// point (t_J, y_J) of the last Jacobian -- touched once or twice per STEP, not per Newton iteration.
// With COLD_GLOBAL they live in a per-CTA slice of the workspace (coalesced, L2-resident: 31 slots x
// 148 x 384 lanes = 14 MB) and shared memory holds nothing but the factors: 600 B per lane instead of
// 928 B, i.e. 12 warps per SM instead of 8.
// With FD (finite-difference Jacobian, jac=None of the reference) the Jacobian cannot be re-materialised
// from (t_J, y_J): it is kept, with the N adaptive column factors of scipy's num_jac, in the per-CTA
// slice of the workspace as well (N^2 + N more global slots).
template <class Prob, bool COLD_GLOBAL = false, bool FD = false>
struct K1Layout {
    static constexpr int N = Prob::N;
    static constexpr bool SELECT = COLD_GLOBAL;     // no scratch slots: permute right-hand sides with selects
    static constexpr int S_LUR = 0;
    static constexpr int S_LCR = N * N;
    static constexpr int S_LCI = 2 * N * N;
    static constexpr int S_B = 3 * N * N;
    static constexpr int HOT = S_B + (SELECT ? 0 : N);
    static constexpr int C_F = 0;
    static constexpr int C_YOLD = C_F + N;
    static constexpr int C_Q = C_YOLD + N;
    static constexpr int C_YJ = C_Q + 3 * N;
    static constexpr int C_TJ = C_YJ + N;
    static constexpr int COLD = C_TJ + 1;
    static constexpr int SHARED_SLOTS = COLD_GLOBAL ? HOT : HOT + COLD;
    static constexpr int G_J = COLD_GLOBAL ? COLD : 0;       // global slots: [cold state] [J, N x N] [column factors, N]
    static constexpr int G_FAC = G_J + N * N;
    static constexpr int GLOBAL_SLOTS = (COLD_GLOBAL ? COLD : 0) + (FD ? N * N + N : 0);
};

// ---- finite-difference Jacobian: scipy.integrate._ivp.common.num_jac (scipy 1.18.1 common.py:268-452,
// called from radauDAE.py:475-480 with threshold = atol), dense branch, one column at a time; the
// columns are scaled by the reciprocal of h (arithmetic contract) -----------------
constexpr double NJ_FACTOR0 = 0x1.0000000000000p-26;       // EPS ** 0.5
constexpr double NJ_DIFF_REJECT = 0x1.6a09e667f3bcdp-46;   // EPS ** 0.875
constexpr double NJ_DIFF_SMALL = 0x1.0000000000000p-39;    // EPS ** 0.75
constexpr double NJ_DIFF_BIG = 0x1.0000000000000p-13;      // EPS ** 0.25
constexpr double NJ_MIN_FACTOR = 0x1.f400000000000p-43;    // 1e3 * EPS
// f(t, y + h e_i) into fn; largest |fn - f| and max(|f|, |fn|) of that row
template <class Prob, int N>
BRDAE_HD void fd_column(const typename Prob::P &prm, double t, const double (&y)[N], const double (&f)[N], int i, double h,
                        double (&fn)[N], double &md, double &scale) {
    double yp[N];
#pragma unroll
    for (int c = 0; c < N; ++c) yp[c] = y[c] + (c == i ? h : 0.0);
    Prob::rhs(prm, t, yp, fn);
    md = fabs(fn[0] - f[0]);
    double fm = f[0], fnm = fn[0];
#pragma unroll
    for (int r = 1; r < N; ++r) {
        const double d = fabs(fn[r] - f[r]);
        const bool g = d > md;
        md = g ? d : md; fm = g ? f[r] : fm; fnm = g ? fn[r] : fnm;
    }
    scale = fmax(fabs(fm), fabs(fnm));
}
// J (global slots gj(G_J ...)) and the column factors (gj(G_FAC ...)) at (t, y) with f = f(t, y)
template <class Prob, class L, class SG>
BRDAE_HD void num_jac_fd(const typename Prob::P &prm, const SolveArgs &a, double t, const double (&y)[Prob::N],
                         const double (&f)[Prob::N], const SG &gj) {
    constexpr int N = Prob::N;
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        double yi = y[0], fi = f[0];
#pragma unroll
        for (int c = 1; c < N; ++c) { yi = (c == i) ? y[c] : yi; fi = (c == i) ? f[c] : fi; }
        const double ys = (fi >= 0 ? 1.0 : -1.0) * fmax(a.atol[i], fabs(yi));
        double fac = gj(L::G_FAC + i);
        double h = (yi + fac * ys) - yi;
        while (h == 0) { fac *= 10; h = (yi + fac * ys) - yi; }
        double fn[N], md, scale;
        fd_column<Prob, N>(prm, t, y, f, i, h, fn, md, scale);
        if (md < NJ_DIFF_REJECT * scale) {
            const double fac_new = 10 * fac;
            const double h_new = (yi + fac_new * ys) - yi;
            double fn2[N], md2, scale2;
            fd_column<Prob, N>(prm, t, y, f, i, h_new, fn2, md2, scale2);
            if (md * scale2 < md2 * scale) {
                fac = fac_new; h = h_new; scale = scale2; md = md2;
#pragma unroll
                for (int r = 0; r < N; ++r) fn[r] = fn2[r];
            }
        }
        const double rh = 1.0 / h;
#pragma unroll
        for (int r = 0; r < N; ++r) gj(L::G_J + r * N + i) = (fn[r] - f[r]) * rh;
        if (md < NJ_DIFF_SMALL * scale) fac *= 10;
        if (md > NJ_DIFF_BIG * scale) fac *= 0.1;
        gj(L::G_FAC + i) = fmax(fac, NJ_MIN_FACTOR);
    }
}

// Per-problem scheduling traits (results do not depend on them; measured in profiles/r01_k1_fastpaths.md):
//   kStaticLU       -- the iteration matrices of this problem practically never need a row exchange: factorise
//                      optimistically on compile-time row addresses (falls back per system on the first off-diagonal pivot);
//   kFirstIterSkip  -- the lanes phase-lock onto the cycle, so whole warps sit at the first Newton iteration together:
//                      skip the rate arithmetic of the convergence test there.
// A problem functor opts in with `static constexpr bool kStaticLU = true;` etc.; the default is off.
template <class P, class = void> struct trait_static_lu { static constexpr bool value = false; };
template <class P> struct trait_static_lu<P, decltype((void)P::kStaticLU)> { static constexpr bool value = P::kStaticLU; };
template <class P, class = void> struct trait_first_iter_skip { static constexpr bool value = false; };
template <class P> struct trait_first_iter_skip<P, decltype((void)P::kFirstIterSkip)> { static constexpr bool value = P::kFirstIterSkip; };

// Event functions of the problem functor itself (plugin builds): `static constexpr int NEVENTS` and
// `static double event(const P &, int k, double t, const double (&y)[N])`; problems without them have none.
template <class P, class = void> struct trait_n_events { static constexpr int value = 0; };
template <class P> struct trait_n_events<P, decltype((void)P::NEVENTS)> { static constexpr int value = P::NEVENTS; };
template <class Prob>
BRDAE_HD double functor_event(const typename Prob::P &prm, int k, double t, const double (&y)[Prob::N]) {
    if constexpr (trait_n_events<Prob>::value > 0) return (k < trait_n_events<Prob>::value) ? Prob::event(prm, k, t, y) : 0.0;
    else return 0.0;
}

// warp-level helpers; the host build (tests/host_sim) runs a "warp" of one lane
#if defined(__CUDA_ARCH__)
BRDAE_D bool warp_all(bool x) { return __all_sync(0xffffffffu, x); }
BRDAE_D bool warp_any(bool x) { return __any_sync(0xffffffffu, x); }
BRDAE_D int warp_count(bool x) { return __popc(__ballot_sync(0xffffffffu, x)); }
BRDAE_D unsigned warp_ballot(bool x) { return __ballot_sync(0xffffffffu, x); }
BRDAE_D unsigned warp_ballot_in(unsigned mask, bool x) { return __ballot_sync(mask, x); }     // among the lanes of `mask` (all of them call)
// Once per cycle: is any lane of the scope still alive?  CTA-wide this is the one barrier of the
// cycle (BAR.RED.OR): it keeps the warps of the SM in step, so that they walk through the same block
// at the same time and share its instruction-cache lines (with free-running warps the kernel was
// instruction-fetch bound: SM I-cache hit rate 51 %, profiles/r01_k1_ncu_summary.md).
template <bool CTA_WIDE>
BRDAE_D bool cycle_sync(bool alive) {
    if (CTA_WIDE) return __syncthreads_or(alive) != 0;
    return __any_sync(0xffffffffu, alive);
}
BRDAE_D void warp_converge() { __syncwarp(); }
BRDAE_D long long fetch_system(bool want, unsigned long long *counter) {
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
// Completion signal of the pipelined copy-out: count this system into its chunk; whoever completes the chunk
// raises its flag in host-mapped memory (the "last block" pattern: every lane fences its results before the
// count, the last one fences system-wide before the flag).
BRDAE_D void signal_chunk(const SolveArgs &a, long long idx) {
    __threadfence();
    const long long k = idx / a.chunk_size, first = k * a.chunk_size;
    const int expect = (int)((a.B - first < a.chunk_size) ? a.B - first : a.chunk_size);
    if (atomicAdd(&a.chunk_done[k], 1) + 1 == expect) {
        __threadfence_system();
        a.chunk_flags[k] = 1;
    }
}
#else
inline bool warp_all(bool x) { return x; }
inline bool warp_any(bool x) { return x; }
inline int warp_count(bool x) { return x ? 1 : 0; }
inline unsigned warp_ballot(bool x) { return x ? 1u : 0u; }
inline unsigned warp_ballot_in(unsigned, bool x) { return x ? 1u : 0u; }
template <bool CTA_WIDE>
inline bool cycle_sync(bool alive) { return alive; }
inline void warp_converge() {}
inline long long fetch_system(bool want, unsigned long long *counter) { return want ? (long long)((*counter)++) : -1; }
#endif

// The per-lane integrator.  `sm` addresses this lane's hot (shared-memory) slots, `cd` its cold ones.
// `t_eval` is the output grid as the lanes should read it (the kernel stages short grids in shared
// memory: the check "is the next output point inside this step" runs after every accepted step, and a
// global load there put an L2 round trip on the critical path of every ERREST pass).
// `gj` addresses this lane's global-memory slots of the finite-difference mode (unused otherwise).
template <class Prob, class Mass, bool CTA_WIDE, bool COLD_GLOBAL, bool FD, bool EVENTS, class S, class SC, class SG>
BRDAE_HD void k1_lane_loop(const SolveArgs &a, const S &sm, const SC &cd, const SG &gj, const double *t_eval) {
    constexpr int N = Prob::N;
    using L = K1Layout<Prob, COLD_GLOBAL, FD>;
    // sparsity of the factors while nothing pivots (dense unless the functor declares jac_nz; finite-difference Jacobians and a
    // user-supplied constant matrix have no declared structure)
    constexpr bool PAT_DENSE = FD || !trait_jac_nz<Prob>::declared;
    using PAT = LuPat<Prob, Mass, PAT_DENSE>;
    const bool pat_ok = PAT_DENSE || !a.has_jac_const;
    typename Prob::P prm;

    // ---- lane state ---------------------------------------------------------------------
    int state = ST_IDLE, setup_kind = SETUP_STEP;
    long long idx = -1;
    double t = 0, y[N];
    double h_abs_state = 0, h_abs_old_state = 0, err_old_state = 0;
    bool hist_state = false, current_jac = true, lu_valid = false, have_sol = false, lu_static_ok = true;
    double inv_h_sol = 1, t_sol_old = 0, inv_h_lu = 1.0;
    unsigned piv_r = 0, piv_c = 0;
    int cnt[CN_COUNT];
    int ei = 0;
    long long nsteps = 0;
    // step scope
    double min_step = 0, h_abs = 0, h_abs_old = 0, err_old = 0;
    bool hist = false, rejected = false;
    // attempt scope
    double h = 0, inv_h = 0, t_new = 0, inv_ns[N], mr = 0, mcr = 0, mci = 0;
    // Newton scope
    double W[3][N];
    int k = 0, nbad = 0, n_iter = 0;
    double dwn_old = 0, rate = 0;
    bool have_old = false, have_rate = false;
#pragma unroll
    for (int c = 0; c < N; ++c) { y[c] = 0; inv_ns[c] = 0; W[0][c] = W[1][c] = W[2][c] = 0; }
#pragma unroll
    for (int q = 0; q < CN_COUNT; ++q) cnt[q] = 0;

    const int MAXIT = a.maxit;
    const double dir = a.dir;

    for (;;) {
        int finish_status = RUNNING;
        // ================= refill idle lanes =================================================
        if (warp_any(state == ST_IDLE)) {
            const long long got = fetch_system(state == ST_IDLE, a.work_counter);
            if (state == ST_IDLE) {
                if (got >= a.B) state = ST_EXIT;
                else {
                    idx = got;
                    bool input_lost = false;
#if defined(__CUDA_ARCH__)
                    if (a.input_ready) {           // pipelined copy-in: this system's chunk may still be on its way
                        const volatile int32_t *ready = a.input_ready + idx / a.chunk_size;
                        // bounded wait: the host raises abort_flag when it abandons the call, and a copy that has not
                        // arrived after ~2^26 polls (tens of seconds) never will (a driver mode in which the launch
                        // blocks the queueing of the copies): give the system up and tell the host
                        unsigned spins = 0;
                        while (*ready == 0) {
                            // (abort_flag is host-mapped: polled once per 65,536 spins, about every 17 ms -- tens of thousands of
                            // waiting lanes reading it on every spin put a storm of PCIe reads in front of the very copies
                            // they wait for: the end-to-end rate fell from 13.6 to 10.3 M trajectories/s)
                            if (((++spins & 0xffffu) == 0 && a.abort_flag && *a.abort_flag != 0) || spins > (1u << 26) ||
                                *reinterpret_cast<volatile unsigned long long *>(a.work_counter + 1) != 0) { input_lost = true; break; }
                            __nanosleep(256);
                        }
                        __threadfence();
                        if (input_lost) atomicExch(a.work_counter + 1, 1ull);
                    }
#endif
                    if (input_lost) {              // never integrated: no result is written, the host returns an error
                        if (a.chunk_done) {
#if defined(__CUDA_ARCH__)
                            signal_chunk(a, idx);
#endif
                        }
                        state = ST_IDLE;
                        idx = -1;
                    } else {
                    Prob::load(prm, a.params, a.par_sc, idx * a.par_sb);
                    t = a.t0;
#pragma unroll
                    for (int c = 0; c < N; ++c) y[c] = a.y0[c * a.y0_sc + idx * a.y0_sb];
                    double f0[N];
                    Prob::rhs(prm, t, y, f0);                            // :304
#pragma unroll
                    for (int c = 0; c < N; ++c) { cd(L::C_F + c) = f0[c]; cd(L::C_YJ + c) = y[c]; }
                    cd(L::C_TJ) = t;                                     // J0 = jac(t0, y0) :483
                    if constexpr (!PAT_DENSE) zero_outside_pattern<PAT, N>(sm, L::S_LUR, L::S_LCR, L::S_LCI);
#pragma unroll
                    for (int q = 0; q < CN_COUNT; ++q) cnt[q] = 0;
                    cnt[CN_NFEV] = 1;
                    cnt[CN_NJEV] = a.has_jac_const ? 0 : 1;
                    h_abs_state = a.h0;
                    hist_state = false; current_jac = true; lu_valid = false; have_sol = false; lu_static_ok = true;
                    inv_h_lu = 1.0; ei = 0; nsteps = 0;
                    state = ST_SETUP; setup_kind = SETUP_STEP;
                    if (EVENTS) {
                        for (int e = 0; e < a.n_events; ++e) a.ev_count[e * a.B + idx] = 0;
                        if (a.rep_cap > 0) a.rep_count[idx] = 0;
                    }
                    if (FD && !a.has_jac_const) {                        // J0 = num_jac(t0, y0, f0) :475-481
#pragma unroll
                        for (int c = 0; c < N; ++c) gj(L::G_FAC + c) = NJ_FACTOR0;
                        state = ST_JACFD;
                    }
                    if (t == a.tb) {  // OdeSolver.step corner case (scipy base.py:193-198)
                        for (; ei < a.T; ++ei)
#pragma unroll
                            for (int c = 0; c < N; ++c) a.y_eval[ei * a.ye_st + c * a.ye_sc + idx * a.ye_sb] = y[c];
                        finish_status = BRDAE_STATUS_FINISHED;
                        state = ST_IDLE;
                    }
                    }
                }
            }
        }
        warp_converge();

        // ================= one barrier per cycle; leave when every lane of the scope is done ==========
        if (!cycle_sync<CTA_WIDE>(state != ST_EXIT)) break;

        // ================= error estimate, accept/reject, epilogue :660-784 ===================
        if (warp_any(state == ST_ERREST)) {
            const bool ident_r = BRDAE_K1_IDENT_FASTPATH && pat_ok && warp_all(state != ST_ERREST || piv_r == perm_identity(N)) && BRDAE_SIM_NEIGHBOUR();
            if (state == ST_ERREST) {
                double Z[3][N];
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    Z[0][c] = stage_z<0>(W[0][c], W[1][c], W[2][c]);
                    Z[1][c] = stage_z<1>(W[0][c], W[1][c], W[2][c]);
                    Z[2][c] = stage_z<2>(W[0][c], W[1][c], W[2][c]);
                }
                const double safety = a.safety_factor * (2 * MAXIT + 1) / (2 * MAXIT + n_iter);   // :631
                double error_norm = 0.0;
                bool accepted = true;
                if (!a.constant_dt) {
                    double ze[N], inv_es[N], mze[N], err[N];
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        double v = Z[0][c] * RE0;
                        v = fma(Z[1][c], RE1, v);
                        v = fma(Z[2][c], RE2, v);
                        ze[c] = v * inv_h;
                        const double ynew = y[c] + Z[2][c];
                        const int ve = a.var_index[c] > 1 ? a.var_index[c] - 1 : 0;
                        const double hp = a.scale_error ? hpow(h, ve) : 1.0;
                        inv_es[c] = hp / (a.atol[c] + fmax(fabs(y[c]), fabs(ynew)) * a.rtol);
                    }
                    mass_mv<Mass>(prm, a, ze, mze);
#pragma unroll
                    for (int c = 0; c < N; ++c) err[c] = cd(L::C_F + c) + mze[c];
                    double err1 = NAN, err2 = NAN;          // per-attempt report only (instrumented instantiation)
                    for (int pass = 0;; ++pass) {
                        if (ident_r) solve_real_ident_p<PAT, N>(sm, L::S_LUR, err);   // rhs NOT scaled by hscale
                        else solve_real_sm<N, L::SELECT>(sm, L::S_LUR, L::S_B, piv_r, err);
                        cnt[CN_NSOLVE_ERR]++;
                        double s = 0.0;
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            if (a.zero_algebraic_error && a.var_index[c] != 0) err[c] = 0.0;
                            else { const double v = err[c] * inv_es[c]; s = fma(v, v, s); }
                        }
                        error_norm = sqrt(s) * (a.zero_algebraic_error ? a.rsq_nondiff : a.rsqn);
                        if (EVENTS) { if (pass == 0) err1 = error_norm; else err2 = error_norm; }
                        if (pass == 1 || !(error_norm > 1 && (a.always_2nd || rejected))) break;
                        double Y[N], F[N];
#pragma unroll
                        for (int c = 0; c < N; ++c) Y[c] = y[c] + err[c];
                        Prob::rhs(prm, t, Y, F);
                        cnt[CN_NFEV]++;
#pragma unroll
                        for (int c = 0; c < N; ++c) err[c] = F[c] + mze[c];
                    }
                    if (EVENTS && a.rep_cap > 0)
                        report_event(a, idx, error_norm > 1 ? RC_REJECTED : RC_ACCEPTED, t, h, n_iter, nbad, err1, err2);   // :719, :736
                    if (error_norm > 1) {                                               // :713-733
                        const double factor = gustafsson(h_abs, h_abs_old, error_norm, err_old, hist, a.predictive_controller);
                        h_abs *= fmax(a.min_factor, safety * factor);
                        lu_valid = false; rejected = true; cnt[CN_NREJ]++;
                        accepted = false;
                        state = ST_SETUP; setup_kind = SETUP_ATTEMPT;
                    } else cnt[CN_NACC]++;
                }
                if (accepted) {                                                         // :742-784
                    const bool has_jac_fn = !a.has_jac_const;
                    const bool recompute = has_jac_fn && n_iter > 2 && have_rate && rate > a.jac_recompute;
                    double factor;
                    if (a.constant_dt) factor = a.max_step / h_abs;
                    else factor = fmin(a.max_factor, safety * gustafsson(h_abs, h_abs_old, error_norm, err_old, hist, a.predictive_controller));
                    if (!recompute && factor < 1.2) factor = 1; else lu_valid = false;
                    double fnew[N];
#pragma unroll
                    for (int c = 0; c < N; ++c) { cd(L::C_YOLD + c) = y[c]; y[c] = y[c] + Z[2][c]; }
                    Prob::rhs(prm, t_new, y, fnew);
                    cnt[CN_NFEV]++;
#pragma unroll
                    for (int c = 0; c < N; ++c) cd(L::C_F + c) = fnew[c];
                    if (recompute) {
#pragma unroll
                        for (int c = 0; c < N; ++c) cd(L::C_YJ + c) = y[c];
                        cd(L::C_TJ) = t_new;
                        cnt[CN_NJEV]++;
                        current_jac = true;
                    } else if (has_jac_fn) current_jac = false;
                    h_abs_old_state = h_abs_state; err_old_state = error_norm; hist_state = true;
                    h_abs_state = h_abs * factor;
                    double Q[N][3];
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        Q[c][0] = fma(Z[2][c], P20, fma(Z[1][c], P10, Z[0][c] * P00));
                        Q[c][1] = fma(Z[2][c], P21, fma(Z[1][c], P11, Z[0][c] * P01));
                        Q[c][2] = fma(Z[2][c], P22, fma(Z[1][c], P12, Z[0][c] * P02));
                        cd(L::C_Q + 3 * c) = Q[c][0]; cd(L::C_Q + 3 * c + 1) = Q[c][1]; cd(L::C_Q + 3 * c + 2) = Q[c][2];
                    }
                    t_sol_old = t; inv_h_sol = inv_h; have_sol = true;
                    t = t_new;
                    bool ev_stop = false;
                    double x_cut = 1.0;
                    if (EVENTS) {
                        // scipy solve_ivp(events=...) on this step (ivp.py find_active_events / handle_events, which
                        // radauDAE.py:1110-1129 wraps): an event function that changed sign in its direction is located on
                        // the step's cubic -- for g linear in y the scalar cubic g_old + a1 x + a2 x^2 + a3 x^3 -- by
                        // bisection to the last bit; occurrences are counted and recorded per event, and the earliest root
                        // whose event reached its terminal count ends the system there (later roots of the step are dropped)
                        double xr[BRDAE_MAX_EVENTS];
                        bool act[BRDAE_MAX_EVENTS];
                        int stop_e = -1;
                        double x_stop = 2.0;
#pragma unroll 1
                        for (int e = 0; e < a.n_events; ++e) {
                            double g0 = -a.ev_level[e], g1 = -a.ev_level[e], a1 = 0.0, a2 = 0.0, a3 = 0.0;
                            // >= 0: a (nonlinear) event function of the problem functor; problems that define none compile to the
                            // linear form alone
                            const int fk = (trait_n_events<Prob>::value > 0) ? a.ev_functor[e] : -1;
                            double yo[N];
#pragma unroll
                            for (int c = 0; c < N; ++c) yo[c] = cd(L::C_YOLD + c);
                            if (fk >= 0) {
                                g0 = functor_event<Prob>(prm, fk, t_sol_old, yo);
                                g1 = functor_event<Prob>(prm, fk, t, y);
                            } else {
#pragma unroll
                                for (int c = 0; c < N; ++c) {
                                    const double ce = a.ev_coef[e][c];
                                    g0 = fma(ce, yo[c], g0); g1 = fma(ce, y[c], g1);
                                    a1 = fma(ce, Q[c][0], a1); a2 = fma(ce, Q[c][1], a2); a3 = fma(ce, Q[c][2], a3);
                                }
                            }
                            const bool up = g0 <= 0 && g1 >= 0, down = g0 >= 0 && g1 <= 0;
                            act[e] = a.ev_dir[e] > 0 ? up : (a.ev_dir[e] < 0 ? down : (up || down));
                            xr[e] = 2.0;
                            if (act[e]) {
                                double lo = 0.0, hi = 1.0, x = 0.0;
                                if (g0 == 0) x = 0.0;
                                else if (g1 == 0) x = 1.0;
                                else {
                                    const bool pos0 = g0 > 0;
                                    for (int it = 0; it < 64 && lo < hi; ++it) {
                                        const double mid = 0.5 * (lo + hi);
                                        if (!(mid > lo && mid < hi)) break;
                                        double pm;
                                        if (fk >= 0) {           // g(t, sol(t)) on the step's interpolant, as scipy's root finder sees it
                                            double ym[N];
                                            const double m2 = mid * mid, m3 = m2 * mid;
#pragma unroll
                                            for (int c = 0; c < N; ++c) ym[c] = fma(Q[c][2], m3, fma(Q[c][1], m2, Q[c][0] * mid)) + yo[c];
                                            pm = functor_event<Prob>(prm, fk, fma(mid, h, t_sol_old), ym);
                                        } else pm = fma(mid, fma(mid, fma(mid, a3, a2), a1), g0);
                                        if ((pm > 0) == pos0 && pm != 0) lo = mid; else hi = mid;
                                    }
                                    x = hi;
                                }
                                xr[e] = x;
                                const int cnt = a.ev_count[e * a.B + idx] + 1;
                                if (a.ev_term[e] > 0 && cnt >= a.ev_term[e] && x < x_stop) { x_stop = x; stop_e = e; }
                            }
                        }
#pragma unroll 1
                        for (int e = 0; e < a.n_events; ++e) {
                            if (!act[e]) continue;
                            if (stop_e >= 0 && !(xr[e] < x_stop || (xr[e] == x_stop && e <= stop_e))) continue;
                            const int k_ev = a.ev_count[e * a.B + idx];
                            a.ev_count[e * a.B + idx] = k_ev + 1;
                            if (k_ev < a.ev_cap) {
                                const double x = xr[e], x2 = x * x, x3 = x2 * x;
                                a.ev_t[((int64_t)e * a.ev_cap + k_ev) * a.B + idx] = fma(x, h, t_sol_old);
                                if (a.ev_y)
#pragma unroll
                                    for (int c = 0; c < N; ++c)
                                        a.ev_y[(((int64_t)e * a.ev_cap + k_ev) * N + c) * a.B + idx] =
                                            fma(Q[c][2], x3, fma(Q[c][1], x2, Q[c][0] * x)) + cd(L::C_YOLD + c);
                            }
                        }
                        ev_stop = stop_e >= 0;
                        x_cut = x_stop;
                    }
                    if (a.dense_cap > 0 && nsteps <= a.dense_cap) {      // OdeSolution record of this step
                        const int64_t ks = nsteps - 1;         // one accepted step per step() call
                        a.dense_t[ks * a.B + idx] = t;
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            a.dense_y[(ks * N + c) * a.B + idx] = y[c];
#pragma unroll
                            for (int j = 0; j < 3; ++j) a.dense_q[((ks * N + c) * 3 + j) * a.B + idx] = Q[c][j];
                        }
                    }
                    if (EVENTS && ev_stop) {     // status 1: the result ends at the event (ivp.py: t = roots[-1]; y = sol(t));
                        const double x = x_cut, x2 = x * x, x3 = x2 * x;     // the step record above keeps the whole step
                        t = fma(x, h, t_sol_old);
#pragma unroll
                        for (int c = 0; c < N; ++c) y[c] = fma(Q[c][2], x3, fma(Q[c][1], x2, Q[c][0] * x)) + cd(L::C_YOLD + c);
                    }
                    // driver: t_eval points inside (t_old, t]  (searchsorted side='right', ivp.py:711-729)
                    while (ei < a.T) {
                        const double te = t_eval[ei];
                        if (!(dir > 0 ? te <= t : te >= t)) break;
                        const double x = (te - t_sol_old) * inv_h_sol;
                        const double x2 = x * x, x3 = x2 * x;
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            double v = Q[c][0] * x;
                            v = fma(Q[c][1], x2, v);
                            v = fma(Q[c][2], x3, v);
                            a.y_eval[ei * a.ye_st + c * a.ye_sc + idx * a.ye_sb] = v + cd(L::C_YOLD + c);
                        }
                        ++ei;
                    }
                    if (EVENTS && ev_stop) finish_status = BRDAE_STATUS_EVENT;
                    else if (dir * (t - a.tb) >= 0) finish_status = BRDAE_STATUS_FINISHED;
                    else { state = (FD && recompute) ? ST_JACFD : ST_SETUP; setup_kind = SETUP_STEP; }
                }
            }
        }
        warp_converge();

        // ================= finite-difference Jacobian at (t, y), f = f(t, y) :475-481 ===========
        if (FD && warp_any(state == ST_JACFD)) {
            if (state == ST_JACFD) {
                double fcur[N];
#pragma unroll
                for (int c = 0; c < N; ++c) fcur[c] = cd(L::C_F + c);
                num_jac_fd<Prob, L>(prm, a, t, y, fcur, gj);
                state = ST_SETUP;
            }
            warp_converge();
        }
        // ================= cheap set-up: prologue :516-562, attempt :564-588, guess :580-583 ===
        if (warp_any(state == ST_SETUP && finish_status == RUNNING)) {
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
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            const int ve = a.var_index[c] > 1 ? a.var_index[c] - 1 : 0;
                            const double hp = a.scale_newton_norm ? hpow(h, ve) : 1.0;
                            inv_ns[c] = hp / (a.atol[c] + fabs(y[c]) * a.rtol);
                        }
                        inv_h = 1.0 / h;   // the one division by h of an attempt
                        mr = MU_REAL * inv_h; mcr = MU_CR * inv_h; mci = MU_CI * inv_h;
                    }
                }
                if (finish_status == RUNNING) {
                    // W = TI Z0, Z0 from the previous step's cubic evaluated beyond its interval
                    if (!have_sol || !a.extrapolate) {
#pragma unroll
                        for (int c = 0; c < N; ++c) { W[0][c] = 0.0; W[1][c] = 0.0; W[2][c] = 0.0; }
                    } else {
                        double x[3], x2[3], x3[3];
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const double tt = t + h * stage_c(i);
                            x[i] = (tt - t_sol_old) * inv_h_sol;
                            x2[i] = x[i] * x[i];
                            x3[i] = x2[i] * x[i];
                        }
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            const double q0 = cd(L::C_Q + 3 * c), q1 = cd(L::C_Q + 3 * c + 1), q2 = cd(L::C_Q + 3 * c + 2);
                            const double yo = cd(L::C_YOLD + c);
                            double z[3];
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                double v = q0 * x[i];
                                v = fma(q1, x2[i], v);
                                v = fma(q2, x3[i], v);
                                v = v + yo;
                                z[i] = v - y[c];
                            }
                            W[0][c] = fma(TI02, z[2], fma(TI01, z[1], TI00 * z[0]));
                            W[1][c] = fma(TI12, z[2], fma(TI11, z[1], TI10 * z[0]));
                            W[2][c] = fma(TI22, z[2], fma(TI21, z[1], TI20 * z[0]));
                        }
                    }
                    k = 0; nbad = 0; have_old = false; have_rate = false; rate = 0; dwn_old = 0;
                    state = lu_valid ? ST_NEWTON : ST_FACTOR;
                }
            }
        }
        // ================= retire a finished / failed system ==================================
        if (warp_any(finish_status != RUNNING)) {
            if (finish_status != RUNNING) {
                const double qnan = NAN;
                for (int e = ei; e < a.T; ++e)
#pragma unroll
                    for (int c = 0; c < N; ++c) a.y_eval[e * a.ye_st + c * a.ye_sc + idx * a.ye_sb] = qnan;
                a.n_eval[idx] = ei;
                a.t_final[idx] = t;
#pragma unroll
                for (int c = 0; c < N; ++c) a.y_final[c * a.yf_sc + idx * a.yf_sb] = y[c];
                a.status[idx] = finish_status;
#pragma unroll
                for (int q = 0; q < CN_COUNT; ++q) a.counters[q * a.cn_sq + idx * a.cn_sb] = cnt[q];
#if defined(__CUDA_ARCH__)
                if (a.chunk_done) signal_chunk(a, idx);
#endif
                state = ST_IDLE;
            }
        }
        // ================= (re)factorisation :594-613 =========================================
        if (warp_any(state == ST_FACTOR)) {
            const bool static_lu = BRDAE_K1_STATIC_LU && trait_static_lu<Prob>::value && pat_ok && warp_all(state != ST_FACTOR || lu_static_ok) && BRDAE_SIM_NEIGHBOUR();
            if (state == ST_FACTOR) {
                if (a.scale_residuals) inv_h_lu = inv_h;
                double J[N][N];
                if (a.has_jac_const) {
#pragma unroll
                    for (int r = 0; r < N; ++r)
#pragma unroll
                        for (int c = 0; c < N; ++c) J[r][c] = a.jac_const[r * N + c];
                } else if (FD) {
#pragma unroll
                    for (int r = 0; r < N; ++r)
#pragma unroll
                        for (int c = 0; c < N; ++c) J[r][c] = gj(L::G_J + r * N + c);
                } else {
                    double yJ[N];
#pragma unroll
                    for (int c = 0; c < N; ++c) yJ[c] = cd(L::C_YJ + c);
                    Prob::jac(prm, cd(L::C_TJ), yJ, J);
                }
                double hs[N];
#pragma unroll
                for (int r = 0; r < N; ++r) hs[r] = (a.scale_residuals && a.var_index[r] >= 1) ? inv_h_lu : 1.0;
                auto col_real = [&](int kc, double (&col)[N]) {
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        const double v = Mass::nz(r, kc) ? fma(mr, Mass::get(prm, a, r, kc), -J[r][kc]) : -J[r][kc];
                        col[r] = hs[r] * v;
                    }
                };
                auto col_cplx = [&](int kc, double (&cr)[N], double (&ci)[N]) {
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        const bool nz = Mass::nz(r, kc);
                        const double m = nz ? Mass::get(prm, a, r, kc) : 0.0;
                        const double v = nz ? fma(mcr, m, -J[r][kc]) : -J[r][kc];
                        cr[r] = hs[r] * v;
                        ci[r] = nz ? hs[r] * (mci * m) : 0.0;
                    }
                };
                // Optimistic factorisation without row exchanges while this system has never needed one (and no lane of
                // the warp would have to run the general code next to it); a pivot off the diagonal repeats the
                // factorisation with the general routine and switches the system over for good.
                bool need_general = !static_lu;
                if (static_lu) {
                    // the same columns on the structural non-zeros of mu M - J (entries outside the pattern: exact zeros)
                    const PatColumns<PAT, Prob, Mass, N> colp{prm, a, J, hs, mr, mcr, mci};
                    const bool ok = lu_real_static_p<PAT, N>(sm, L::S_LUR, colp) && lu_cplx_static_p<PAT, N>(sm, L::S_LCR, L::S_LCI, colp);
                    piv_r = piv_c = perm_identity(N);
                    BRDAE_SIM_COUNT(0);
                    if (!ok) { need_general = true; lu_static_ok = false; BRDAE_SIM_COUNT(1); }
                }
                if (need_general) {
                    BRDAE_SIM_COUNT(2);
                    piv_r = lu_real_left<N>(sm, L::S_LUR, col_real);
                    piv_c = lu_cplx_left<N>(sm, L::S_LCR, L::S_LCI, col_cplx);
                    // A system that exchanged rows here -- also when a NEIGHBOUR forced the general routine on the warp -- stays
                    // on it: the exchange leaves fill outside the no-exchange pattern, which the pattern-restricted optimistic
                    // factorisation would not overwrite and a later general substitution (forced by a neighbour again) would read.
                    if (piv_r != perm_identity(N) || piv_c != perm_identity(N)) lu_static_ok = false;
                }
                cnt[CN_NLU]++;
                if (EVENTS && a.rep_cap > 0) report_event(a, idx, RC_FACTORISATION, t, h, -1, -1, -1.0, -1.0);   // :607
                lu_valid = true;
                state = ST_NEWTON;
            }
        }
        warp_converge();
        // ================= simplified-Newton iterations :846-940 ==============================
        // a rolled loop (one copy of the code) of a.newton_reps passes; a lane that converged or failed
        // sits the remaining passes out
#pragma unroll 1
        for (int rep = 0; rep < a.newton_reps; ++rep) {
        if (warp_any(state == ST_NEWTON)) {
            // no lane of the warp exchanged a row in its current factorisations (always, for the pendulum and Robertson
            // ensembles): the substitutions below run on compile-time addresses, without the permutation round trip
            const unsigned perm_id = perm_identity(N);
            const bool ident = BRDAE_K1_IDENT_FASTPATH && pat_ok && warp_all(state != ST_NEWTON || (piv_r == perm_id && piv_c == perm_id)) && BRDAE_SIM_NEIGHBOUR();
            // every Newton lane of the warp is at the first iteration of its attempt (the lanes phase-lock onto the cycle, so
            // this is the common case in the first pass): there is no previous increment norm, hence no rate, and the
            // division / power / reciprocal chain of the convergence test would only produce values nobody reads
            const bool any_old = !(BRDAE_K1_FIRST_ITER_FASTPATH && trait_first_iter_skip<Prob>::value) || warp_any(state == ST_NEWTON && have_old);
            if (state == ST_NEWTON) {
                // the three stages as one loop body with table coefficients (unrolled by default).  Values are those of the unrolled form: the first
                // accumulation fma(a, F, 0) equals a*F, and (T W)_2 = fma(0, W2, fma(1, W1, 1*W0)) = W0 + W1.
                double fr[N], fcr[N], fci[N];
                bool finite = true;
#pragma unroll
                for (int c = 0; c < N; ++c) { fr[c] = 0.0; fcr[c] = 0.0; fci[c] = 0.0; }
#if defined(BRDAE_ROLL_STAGES)   // smaller code; measured 1.6 % slower once all warps of an SM share the block
#pragma unroll 1
#else
#pragma unroll
#endif
                for (int i = 0; i < 3; ++i) {
                    const double t0c = kStageT[i][0], t1c = kStageT[i][1], t2c = kStageT[i][2];
                    const double a0 = kStageTI[0][i], a1 = kStageTI[1][i], a2 = kStageTI[2][i];
                    double Y[N], F[N];
#pragma unroll
                    for (int c = 0; c < N; ++c) Y[c] = y[c] + fma(t2c, W[2][c], fma(t1c, W[1][c], t0c * W[0][c]));
                    Prob::rhs(prm, t + h * kStageC[i], Y, F);
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        const double Fv = F[c];
                        finite = finite && is_finite(Fv);
                        fr[c] = fma(a0, Fv, fr[c]); fcr[c] = fma(a1, Fv, fcr[c]); fci[c] = fma(a2, Fv, fci[c]);
                    }
                }
                cnt[CN_NFEV] += 3;
                int outcome = 0;  // 0 continue, 1 converged, 2 stop (not converged)
                if (!finite) outcome = 2;
                else {
                    double mw0[N], mw1[N], mw2[N];
                    mass_mv<Mass>(prm, a, W[0], mw0);
                    mass_mv<Mass>(prm, a, W[1], mw1);
                    mass_mv<Mass>(prm, a, W[2], mw2);
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        const double hs = (a.scale_residuals && a.var_index[c] >= 1) ? inv_h_lu : 1.0;
                        double ar = fr[c], br = fcr[c], bi = fci[c];
                        if (!mass_row_empty<Mass, N>(c)) {
                            ar = fma(-mr, mw0[c], ar);
                            br = fma(-mcr, mw1[c], br); br = fma(mci, mw2[c], br);
                            bi = fma(-mcr, mw2[c], bi); bi = fma(-mci, mw1[c], bi);
                        }
                        fr[c] = ar * hs; fcr[c] = br * hs; fci[c] = bi * hs;
                    }
                    if (ident) {
                        BRDAE_SIM_COUNT(3);
                        solve_real_ident_p<PAT, N>(sm, L::S_LUR, fr);
                        solve_cplx_ident_p<PAT, N>(sm, L::S_LCR, L::S_LCI, fcr, fci);
                    } else {
                        BRDAE_SIM_COUNT(4);
                        solve_real_sm<N, L::SELECT>(sm, L::S_LUR, L::S_B, piv_r, fr);
                        solve_cplx_sm<N, L::SELECT>(sm, L::S_LCR, L::S_LCI, L::S_B, piv_c, fcr, fci);
                    }
                    cnt[CN_NSOLVE]++;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < N; ++c) { const double v = fr[c] * inv_ns[c]; s = fma(v, v, s); }
#pragma unroll
                    for (int c = 0; c < N; ++c) { const double v = fcr[c] * inv_ns[c]; s = fma(v, v, s); }
#pragma unroll
                    for (int c = 0; c < N; ++c) { const double v = fci[c] * inv_ns[c]; s = fma(v, v, s); }
                    const double dwn = sqrt(s) * a.rsq3n;
#pragma unroll
                    for (int c = 0; c < N; ++c) { W[0][c] += fr[c]; W[1][c] += fcr[c]; W[2][c] += fci[c]; }
                    // convergence logic :880-932 as straight-line predicate algebra: the nested ifs of the
                    // reference compiled to a cascade of divergent branches at the very end of the block's
                    // dependency chain, where nothing is left to hide their latency
                    if (!any_old) {       // first iteration everywhere: rate, have_rate and nbad keep their values (:880-932 with
                        const bool conv_now = dwn < a.newton_tol;      // dW_norm_old is None)
                        BRDAE_SIM_COUNT(5);
                        outcome = conv_now ? 1 : 0;
                        dwn_old = conv_now ? dwn_old : dwn;
                        have_old = !conv_now;
                    } else {
                    const bool had_old = have_old;
                    const double r_new = dwn / (had_old ? dwn_old : 1.0);
                    rate = had_old ? r_new : rate;
                    have_rate = have_rate || had_old;
                    const double inv1 = 1.0 / (1.0 - rate);
                    const double dw_true = rate * inv1 * dwn;
                    double rp = rate;              // rate**(MAXIT-k) by repeated multiplication: straight-line selects for
                    const int npow = MAXIT - k;    // the usual iteration limits, a (uniform) loop beyond
#pragma unroll
                    for (int e = 1; e < 8; ++e) rp = (e < npow) ? rp * rate : rp;
                    for (int e = 8; e < MAXIT; ++e) rp = (e < npow) ? rp * rate : rp;
                    const double dw_true_max = rp * inv1 * dwn;
                    const bool conv_now = dwn < a.newton_tol;
                    const bool live = have_rate && !conv_now;
                    const bool bad_ok = nbad < a.nmax_bad;
                    const bool ge1 = live && rate >= 1;
                    const bool retry_a = ge1 && rate < 100 && bad_ok;                 // diverging, but allowed a bad iteration
                    const bool stop_b = ge1 && !retry_a && !a.all_newton;
                    const bool fall = live && !retry_a && !stop_b;
                    const bool conv_c = fall && dw_true < a.newton_tol;
                    const bool slow_d = fall && !conv_c && dw_true_max > a.newton_tol && a.predictive_newton;
                    const bool retry_d = slow_d && bad_ok, stop_d = slow_d && !bad_ok && !a.all_newton;
                    outcome = (conv_now || conv_c) ? 1 : ((stop_b || stop_d) ? 2 : 0);
                    nbad += (retry_a || retry_d) ? 1 : 0;
                    const bool keep_old = retry_a || retry_d || conv_now || conv_c || stop_b || stop_d;   // :928 dW_norm_old
                    dwn_old = keep_old ? dwn_old : dwn;
                    have_old = have_old || !keep_old;
                    }
                }
                if (outcome == 0) {
                    ++k;
                    if (k >= MAXIT) { outcome = 2; n_iter = MAXIT; }
                } else n_iter = k + 1;
                if (outcome != 0) {
                    if (outcome == 1) state = ST_ERREST;
                    else if (current_jac) {                                            // :650-658
                        if (EVENTS && a.rep_cap > 0) report_event(a, idx, RC_NON_CONV, t, h, n_iter, nbad, -1.0, -1.0);   // :654
                        cnt[CN_NFAIL]++;
                        h_abs *= a.nonconv_factor;
                        lu_valid = false;
                        state = ST_SETUP; setup_kind = SETUP_ATTEMPT;
                    } else {                                                            // :643-647
#pragma unroll
                        for (int c = 0; c < N; ++c) cd(L::C_YJ + c) = y[c];
                        cd(L::C_TJ) = t;
                        cnt[CN_NJEV]++;
                        if (EVENTS && a.rep_cap > 0) report_event(a, idx, RC_JACOBIAN_UPDATE, t, h, -1, -1, -1.0, -1.0);   // :644
                        current_jac = true; lu_valid = false;
                        state = FD ? ST_JACFD : ST_SETUP; setup_kind = SETUP_GUESS;
                    }
                }
            }
        }
        warp_converge();
        }
    }
}

#if defined(__CUDACC__)
template <class Prob, class Mass, int BLOCK, int MIN_BLOCKS, bool CTA_WIDE, bool COLD_GLOBAL, bool FD, bool EVENTS = false>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k1_kernel(const __grid_constant__ SolveArgs a) {
    extern __shared__ double k1_smem[];
    using L = K1Layout<Prob, COLD_GLOBAL, FD>;
    Slots<BLOCK> sm{k1_smem + threadIdx.x};
    Slots<BLOCK> gj{a.scratch + (size_t)blockIdx.x * (L::GLOBAL_SLOTS * BLOCK) + threadIdx.x};   // [slot][lane] slice of this CTA
    Slots<BLOCK> cd{COLD_GLOBAL ? gj.base : k1_smem + L::HOT * BLOCK + threadIdx.x};
    __shared__ double t_eval_sh[K1_TEVAL_SHARED];
    const bool staged = a.T <= K1_TEVAL_SHARED;
    if (staged) {
        for (int i = threadIdx.x; i < a.T; i += BLOCK) t_eval_sh[i] = a.t_eval[i];
        __syncthreads();
    }
    k1_lane_loop<Prob, Mass, CTA_WIDE, COLD_GLOBAL, FD, EVENTS>(a, sm, cd, gj, staged ? t_eval_sh : a.t_eval);
}
#endif

}  // namespace brdae
