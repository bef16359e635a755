// ============================================================================
// oracle/ref_lbfgs.cpp — thin C wrapper around the REFERENCE's own L-BFGS.
// TEST INFRASTRUCTURE ONLY.  `#include "gcopter/lbfgs.hpp"` resolves to the
// unmodified file under /root/reference/src/planner/include (passed with -I by
// oracle/Makefile); <Eigen/Eigen> resolves to oracle/eigen_shim.  The output
// goes to oracle/_ref/libref_lbfgs.so (git-ignored, travels to the GPU box).
// Used to pin oracle/lbfgs_oracle.hpp and, through it, the device driver.
// ============================================================================
#include <cstdint>

#include "gcopter/lbfgs.hpp"

extern "C" {

typedef double (*raw_eval_fn)(void *instance, const double *x, double *g, int n);

// Field order of include/mincob.h:mincob_params / oracle_capi.cpp:orc_params.
struct ref_params {
    int32_t S, kappa;
    double mu, w_pos, w_vel, w_acc, w_jerk, v_max, a_max, j_max, rho;
    int32_t mem_size, past, max_iterations, max_linesearch;
    double g_epsilon, delta, min_step, max_step, f_dec_coeff, s_curv_coeff, cautious_factor, machine_prec;
    int32_t flags, mapping;
};

struct Bridge {
    raw_eval_fn fn;
    void *inst;
    int evals;
    int iters;
};

static double bridge_eval(void *p, const Eigen::VectorXd &x, Eigen::VectorXd &g) {
    Bridge *b = static_cast<Bridge *>(p);
    ++b->evals;
    return b->fn(b->inst, x.data(), g.data(), x.size());
}
static int bridge_progress(void *p, const Eigen::VectorXd &, const Eigen::VectorXd &, const double,
                           const double, const int k, const int) {
    static_cast<Bridge *>(p)->iters = k;
    return 0;
}

int ref_lbfgs_optimize(int n, double *x, double *f, raw_eval_fn eval, void *inst, const ref_params *p,
                       int *iters, int *evals) {
    lbfgs::lbfgs_parameter_t lp;
    lp.mem_size = p->mem_size; lp.g_epsilon = p->g_epsilon; lp.past = p->past; lp.delta = p->delta;
    lp.max_iterations = p->max_iterations; lp.max_linesearch = p->max_linesearch;
    lp.min_step = p->min_step; lp.max_step = p->max_step; lp.f_dec_coeff = p->f_dec_coeff;
    lp.s_curv_coeff = p->s_curv_coeff; lp.cautious_factor = p->cautious_factor;
    lp.machine_prec = p->machine_prec;
    Eigen::VectorXd xv(n);
    for (int i = 0; i < n; ++i) xv(i) = x[i];
    Bridge br{eval, inst, 0, 0};
    double fx = 0.0;
    const int ret = lbfgs::lbfgs_optimize(xv, fx, &bridge_eval, nullptr, &bridge_progress, &br, lp);
    for (int i = 0; i < n; ++i) x[i] = xv(i);
    *f = fx;
    if (iters) *iters = br.iters;   // last k reported by the progress hook
    if (evals) *evals = br.evals;
    return ret;
}

const char *ref_lbfgs_strerror(int code) { return lbfgs::lbfgs_strerror(code); }

}  // extern "C"
