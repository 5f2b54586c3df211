// host_args.hpp -- host-side marshalling shared by the C-ABI translation unit (and by the
// CPU simulation harness under tests/host_sim that single-steps the kernel source for debugging):
// batch validation and the uniform quantities the reference's constructor derives
// (radauDAE.py:263-422: newton_tol :322, NMAX_BAD :400-410, h_abs :307-315).
#pragma once
#include <cmath>
#include <cstring>

#include "radau_common.cuh"

namespace brdae {

constexpr size_t kCounterBytes = 256;

struct ProblemInfo { const char *name; int32_t id, n, n_params; };
static const ProblemInfo kProblems[] = {
    {"pendulum3", BRDAE_PENDULUM3, 5, 3}, {"pendulum2", BRDAE_PENDULUM2, 5, 3},
    {"pendulum1", BRDAE_PENDULUM1, 5, 3}, {"robertson", BRDAE_ROBERTSON, 3, 3},
    {"transistor", BRDAE_TRANSISTOR, 5, 12}, {"npendulum", BRDAE_NPENDULUM, 0, 0},
    {"robertson_ode", BRDAE_ROBERTSON_ODE, 3, 3}, {"jay", BRDAE_JAY, 5, 1},
    {"polydae2", BRDAE_POLYDAE2, 2, 5}, {"linear7", BRDAE_LINEAR7, 7, 7},
};

// the problem of a plugin build: thread-per-system functor (k1_user.cu) or band problem (k2_user.cu); nullptr in the stock library
const ProblemInfo *user_problem_info();
const ProblemInfo *user_band_problem_info();
// problems whose state does not fit a lane: CTA-per-system / lane-per-rope kernels, structure-of-arrays batches only
inline bool is_large_problem(int32_t id) { return id == BRDAE_NPENDULUM || (id == BRDAE_USER && user_band_problem_info() != nullptr); }

inline const ProblemInfo *find_problem(int32_t id) {
    for (const auto &p : kProblems) if (p.id == id) return &p;
    if (id == BRDAE_USER) return user_band_problem_info() ? user_band_problem_info() : user_problem_info();
    return nullptr;
}

inline int validate(const brdae_batch *b, const brdae_options *o) {
    if (!b || !o) return BRDAE_ERR_BAD_ARGUMENT;
    const ProblemInfo *pi = find_problem(b->problem);
    if (!pi) return BRDAE_ERR_UNKNOWN_PROBLEM;
    if (b->n_systems < 0 || b->n_t_eval < 0) return BRDAE_ERR_BAD_ARGUMENT;
    if (pi->n > 0 && b->n != pi->n) return BRDAE_ERR_BAD_ARGUMENT;
    if (pi->n == 0 && (b->n <= 0 || b->n % 5 != 0)) return BRDAE_ERR_BAD_ARGUMENT;
    if (b->n_params != pi->n_params) return BRDAE_ERR_BAD_ARGUMENT;
    if (b->n_systems > 0) {
        if (!b->y0 || !b->n_eval || !b->t_final || !b->y_final || !b->status || !b->counters) return BRDAE_ERR_BAD_ARGUMENT;
        if (pi->n_params > 0 && !b->params) return BRDAE_ERR_BAD_ARGUMENT;
        if (b->n_t_eval > 0 && (!b->t_eval || !b->y_eval)) return BRDAE_ERR_BAD_ARGUMENT;
    }
    if (!b->atol) return BRDAE_ERR_BAD_ARGUMENT;
    if (b->dense_cap < 0 || (b->dense_cap > 0 && !(b->dense_t && b->dense_y && b->dense_q))) return BRDAE_ERR_BAD_ARGUMENT;
    if (b->n_events < 0 || b->n_events > BRDAE_MAX_EVENTS) return BRDAE_ERR_BAD_ARGUMENT;
    if (b->n_events > 0) {
        if (is_large_problem(b->problem)) return BRDAE_ERR_UNSUPPORTED;       // thread-per-system kernels only
        if (b->event_cap < 1 || !b->event_coef || !b->event_level || !b->event_direction || !b->event_terminal) return BRDAE_ERR_BAD_ARGUMENT;
        if (b->n_systems > 0 && (!b->event_count || !b->event_t)) return BRDAE_ERR_BAD_ARGUMENT;
        for (int e = 0; e < b->n_events; ++e) {
            if (b->event_terminal[e] < 0) return BRDAE_ERR_BAD_ARGUMENT;
            if (b->event_functor && b->event_functor[e] >= 0 && b->problem != BRDAE_USER) return BRDAE_ERR_UNSUPPORTED;   // the stock functors define none
        }
    }
    if (b->report_cap < 0 || (b->report_cap > 0 && !(b->report_count && b->report_code && b->report_it && b->report_val))) return BRDAE_ERR_BAD_ARGUMENT;
    if (!(b->rtol > 0) || !(o->max_step > 0) || o->max_newton_ite < 1) return BRDAE_ERR_BAD_ARGUMENT;
    for (int c = 0; c < b->n; ++c) {
        if (!(b->atol[c] >= 0)) return BRDAE_ERR_BAD_ARGUMENT;
        if (b->var_index && (b->var_index[c] < 0 || b->var_index[c] > 3)) return BRDAE_ERR_BAD_ARGUMENT;
    }
    if (o->first_step > 0 && o->first_step > std::fabs(b->t_bound - b->t0)) return BRDAE_ERR_BAD_ARGUMENT;
    return BRDAE_OK;
}

// uniform quantities the reference's constructor derives
inline void fill_common(SolveArgs &a, const brdae_batch *b, const brdae_options *o, void *workspace) {
    std::memset(&a, 0, sizeof(a));
    a.B = b->n_systems; a.T = b->n_t_eval; a.n = b->n;
    a.y0 = b->y0; a.params = b->params; a.t_eval = b->t_eval;
    a.y_eval = b->y_eval; a.n_eval = b->n_eval; a.t_final = b->t_final; a.y_final = b->y_final;
    a.status = b->status; a.counters = b->counters;
    a.y0_sc = a.par_sc = a.yf_sc = a.cn_sq = a.ye_sc = b->n_systems;      // structure-of-arrays
    a.y0_sb = a.par_sb = a.yf_sb = a.cn_sb = a.ye_sb = 1;
    a.ye_st = (int64_t)b->n * b->n_systems;
    a.chunk_done = nullptr; a.chunk_flags = nullptr; a.chunk_size = 0; a.input_ready = nullptr;
    a.dense_t = b->dense_t; a.dense_y = b->dense_y; a.dense_q = b->dense_q;
    a.dense_cap = (b->dense_t && b->dense_y && b->dense_q) ? b->dense_cap : 0;
    a.n_events = b->n_events; a.ev_cap = b->event_cap;
    a.ev_count = b->event_count; a.ev_t = b->event_t; a.ev_y = b->event_y;
    for (int e = 0; e < b->n_events; ++e) {
        a.ev_dir[e] = b->event_direction[e]; a.ev_term[e] = b->event_terminal[e]; a.ev_level[e] = b->event_level[e];
        a.ev_functor[e] = b->event_functor ? b->event_functor[e] : -1;
        for (int c = 0; c < b->n && c < BRDAE_NMAX_SMALL; ++c) a.ev_coef[e][c] = b->event_coef[e * b->n + c];
    }
    a.rep_cap = b->report_cap; a.rep_count = b->report_count; a.rep_code = b->report_code; a.rep_it = b->report_it; a.rep_val = b->report_val;
    a.work_counter = static_cast<unsigned long long *>(workspace);
    a.scratch = reinterpret_cast<double *>(static_cast<char *>(workspace) + kCounterBytes);
    a.t0 = b->t0; a.tb = b->t_bound;
    a.dir = (b->t_bound != b->t0) ? (b->t_bound > b->t0 ? 1.0 : -1.0) : 1.0;     // scipy base.py:166
    a.rtol = b->rtol;
    a.h0 = (o->first_step > 0) ? o->first_step : a.dir * std::fmin(1e-6, std::fabs(b->t_bound - b->t0));  // :313
    a.max_step = o->max_step;
    a.newton_tol = (o->newton_tol > 0) ? o->newton_tol
                   : std::fmax(10 * 2.220446049250313e-16 / b->rtol, std::fmin(0.03, std::sqrt(b->rtol)));  // :322
    a.nonconv_factor = o->factor_on_non_convergence; a.safety_factor = o->safety_factor;
    a.min_factor = o->min_factor; a.max_factor = o->max_factor; a.jac_recompute = o->jac_recompute_factor;
    a.nmax_step = o->nmax_step; a.maxit = o->max_newton_ite;
    int nmax_bad = o->max_bad_ite, n_alg = 0;
    if (nmax_bad < 0) {                                                            // :400-410
        nmax_bad = 0;
        for (int c = 0; c < b->n; ++c) if (b->var_index && b->var_index[c] > 1) nmax_bad = 1;
    }
    for (int c = 0; c < b->n; ++c) if (b->var_index && b->var_index[c] != 0) ++n_alg;
    a.nmax_bad = nmax_bad;
    a.rsq3n = 1.0 / std::sqrt(3.0 * b->n);
    a.rsqn = 1.0 / std::sqrt((double)b->n);
    a.rsq_nondiff = 1.0 / std::sqrt((double)(b->n - n_alg));
    a.scale_residuals = o->scale_residuals; a.scale_newton_norm = o->scale_newton_norm;
    a.scale_error = o->scale_error; a.zero_algebraic_error = o->zero_algebraic_error;
    a.always_2nd = o->always_2nd_estimate; a.predictive_controller = o->predictive_controller;
    a.predictive_newton = o->predictive_newton_stop; a.extrapolate = o->extrapolated_guess;
    a.all_newton = o->all_newton_iterations; a.constant_dt = o->constant_dt;
    a.has_jac_const = b->jac_const != nullptr;
    a.newton_reps = 3;
    a.newton_vote = 0;
    a.jac_fd = (o->jac_finite_differences && !b->jac_const) ? 1 : 0;
    if (b->n <= BRDAE_NMAX_SMALL) {
        for (int c = 0; c < b->n; ++c) {
            a.atol[c] = b->atol[c];
            a.var_index[c] = b->var_index ? b->var_index[c] : 0;
        }
        if (b->mass) std::memcpy(a.mass, b->mass, sizeof(double) * b->n * b->n);
        if (b->jac_const) std::memcpy(a.jac_const, b->jac_const, sizeof(double) * b->n * b->n);
    }
}


}  // namespace brdae
