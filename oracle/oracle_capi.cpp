// ============================================================================
// oracle/oracle_capi.cpp — C entry points (ctypes) over the CPU oracle.
// TEST INFRASTRUCTURE ONLY (see oracle/minco_oracle.hpp header): used by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
// Build: make -C oracle   ->  oracle/liboracle.so
// ============================================================================
#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

#include "lbfgs_oracle.hpp"
#include "minco_oracle.hpp"

extern "C" {

// Same field order as include/mincob.h:mincob_params (tests assert the sizes agree).
struct orc_params {
    int32_t S, kappa;
    double mu, w_pos, w_vel, w_acc, w_jerk, v_max, a_max, j_max, rho;
    int32_t mem_size, past, max_iterations, max_linesearch;
    double g_epsilon, delta, min_step, max_step, f_dec_coeff, s_curv_coeff, cautious_factor, machine_prec;
    int32_t flags, mapping;   // flags: 1 = freeze times, 2 = planner rows [n, b]; mapping: device-only, ignored here
};

int orc_params_size() { return static_cast<int>(sizeof(orc_params)); }

}  // extern "C"

namespace {

orc::PenaltyParams penalty_of(const orc_params &p) {
    orc::PenaltyParams q;
    q.kappa = p.kappa; q.mu = p.mu;
    q.w_pos = p.w_pos; q.w_vel = p.w_vel; q.w_acc = p.w_acc; q.w_jerk = p.w_jerk;
    q.v_max = p.v_max; q.a_max = p.a_max; q.j_max = p.j_max; q.rho = p.rho;
    q.freeze_times = (p.flags & 1) != 0; q.planner_rows = (p.flags & 2) != 0;
    return q;
}
orc::LbfgsParams lbfgs_of(const orc_params &p) {
    orc::LbfgsParams q;
    q.mem_size = p.mem_size; q.g_epsilon = p.g_epsilon; q.past = p.past; q.delta = p.delta;
    q.max_iterations = p.max_iterations; q.max_linesearch = p.max_linesearch;
    q.min_step = p.min_step; q.max_step = p.max_step; q.f_dec_coeff = p.f_dec_coeff;
    q.s_curv_coeff = p.s_curv_coeff; q.cautious_factor = p.cautious_factor; q.machine_prec = p.machine_prec;
    return q;
}

template <int S>
void minco_forward(int N, const double *head, const double *tail, const double *inPs, const double *ts,
                   double *coeffs, double *energy, double *gdC, double *gdT, double *flat) {
    orc::Minco<S> m;
    m.setConditions(head, tail, N);
    m.setParameters(inPs, ts);
    const auto &b = m.getCoeffs();
    if (coeffs) std::copy(b.begin(), b.end(), coeffs);
    if (energy) m.getEnergy(*energy);
    if (gdC) m.getEnergyPartialGradByCoeffs(gdC);
    if (gdT) m.getEnergyPartialGradByTimes(gdT);
    if (flat) m.getTrajectoryFlat(flat);
}
template <int S>
void minco_propagate(int N, const double *head, const double *tail, const double *inPs, const double *ts,
                     const double *gdC, const double *gdT, double *gq, double *gT) {
    orc::Minco<S> m;
    m.setConditions(head, tail, N);
    m.setParameters(inPs, ts);
    m.propogateGrad(gdC, gdT, gq, gT);
}

struct CostBase {
    virtual ~CostBase() = default;
    virtual double eval(const double *x, double *g) = 0;
    virtual int nvars() const = 0;
    virtual void flat(const double *x, double *out, double *T) = 0;
    virtual long evals() const = 0;
};
template <int S>
struct CostImpl : CostBase {
    orc::CostFunctional<S> cf;
    double eval(const double *x, double *g) override { return cf.eval(x, g); }
    int nvars() const override { return cf.nvars(); }
    long evals() const override { return cf.evals; }
    void flat(const double *x, double *out, double *T) override {
        const int N = cf.pb.N;
        std::vector<double> tt(N);
        for (int i = 0; i < N; ++i) tt[i] = orc::forward_t(x[i]);
        cf.minco.setParameters(x + N, tt.data());
        if (out) cf.minco.getTrajectoryFlat(out);
        if (T) std::copy(tt.begin(), tt.end(), T);
    }
};

CostBase *make_cost(const orc_params &p, int N, const double *head, const double *tail,
                    const double *hpolys, const int *hrows, int K) {
    orc::Problem pb;
    pb.N = N; pb.head = head; pb.tail = tail; pb.hpolys = hpolys; pb.hrows = hrows; pb.Kstride = K;
    if (p.S == 3) { auto *c = new CostImpl<3>; c->cf.setup(penalty_of(p), pb); return c; }
    if (p.S == 4) { auto *c = new CostImpl<4>; c->cf.setup(penalty_of(p), pb); return c; }
    return nullptr;
}

double cost_thunk(void *inst, const double *x, double *g, int) { return static_cast<CostBase *>(inst)->eval(x, g); }

}  // namespace

extern "C" {

// ---- MINCO pieces (single problem) ----------------------------------------
int orc_minco_forward(int S, int N, const double *head, const double *tail, const double *inPs,
                      const double *ts, double *coeffs, double *energy, double *gdC, double *gdT, double *flat) {
    if (S == 3) minco_forward<3>(N, head, tail, inPs, ts, coeffs, energy, gdC, gdT, flat);
    else if (S == 4) minco_forward<4>(N, head, tail, inPs, ts, coeffs, energy, gdC, gdT, flat);
    else return -1;
    return 0;
}
int orc_minco_propagate(int S, int N, const double *head, const double *tail, const double *inPs,
                        const double *ts, const double *gdC, const double *gdT, double *gq, double *gT) {
    if (S == 3) minco_propagate<3>(N, head, tail, inPs, ts, gdC, gdT, gq, gT);
    else if (S == 4) minco_propagate<4>(N, head, tail, inPs, ts, gdC, gdT, gq, gT);
    else return -1;
    return 0;
}
// Dense (row-major n x n) -> band LU -> solve (adj=0) or transpose solve (adj=1); b is n x m row-major.
int orc_banded_solve(int n, int p, int q, const double *dense, double *b, int m, int adj) {
    orc::Banded A;
    A.create(n, p, q);
    for (int i = 0; i < n; ++i)
        for (int j = std::max(0, i - p); j <= std::min(n - 1, i + q); ++j) A.at(i, j) = dense[i * n + j];
    A.factorizeLU();
    if (adj) A.solveAdj(b, m); else A.solve(b, m);
    return 0;
}
int orc_smoothed_l1(double mu, double x, double *f, double *df) {
    *f = 0.0; *df = 0.0;
    return orc::smoothed_l1(mu, x, *f, *df) ? 1 : 0;
}
double orc_forward_t(double tau) { return orc::forward_t(tau); }
double orc_backward_t(double T) { return orc::backward_t(T); }
double orc_backward_grad_t(double tau, double gT) { return orc::backward_grad_t(tau, gT); }

// ---- cost functional as an opaque instance + lbfgs-compatible thunk --------
void *orc_cost_create(const orc_params *p, int N, const double *head, const double *tail,
                      const double *hpolys, const int *hrows, int K) {
    return make_cost(*p, N, head, tail, hpolys, hrows, K);
}
void orc_cost_destroy(void *inst) { delete static_cast<CostBase *>(inst); }
double orc_cost_thunk(void *inst, const double *x, double *g, int n) { return cost_thunk(inst, x, g, n); }
void *orc_cost_thunk_address() { return reinterpret_cast<void *>(&orc_cost_thunk); }
void orc_cost_flat(void *inst, const double *x, double *flat, double *T) { static_cast<CostBase *>(inst)->flat(x, flat, T); }

// ---- generic L-BFGS (any callback) -----------------------------------------
int orc_lbfgs_optimize(int n, double *x, double *f, orc::eval_fn eval, void *inst, const orc_params *p,
                       int *iters, int *evals) {
    orc::LbfgsTrace tr;
    const int ret = orc::lbfgs_optimize(n, x, *f, eval, inst, lbfgs_of(*p), &tr);
    if (iters) *iters = tr.iterations;
    if (evals) *evals = tr.evaluations;
    return ret;
}

// ---- batched cost evaluation / optimisation over host threads ---------------
// Layouts as include/mincob.h: head,tail [B][S][3]; hpolys [B][N][K][4]; hrows [B][N];
// x,g [B][n], n = N + 3(N-1); coeffs [B][N][3][2S] (Trajectory order); T [B][N].
static void run_threads(int B, int nthreads, const std::function<void(int, int)> &body);

int orc_cost_batch(const orc_params *p, int B, int N, const double *head, const double *tail,
                   const double *hpolys, const int *hrows, int K, const double *x, double *f, double *g,
                   int nthreads) {
    const int S = p->S, n = N + 3 * (N - 1);
    if (S != 3 && S != 4) return -1;
    run_threads(B, nthreads, [&](int lo, int hi) {
        for (int b = lo; b < hi; ++b) {
            CostBase *c = make_cost(*p, N, head + (size_t)b * 3 * S, tail + (size_t)b * 3 * S,
                                    hpolys ? hpolys + (size_t)b * N * K * 4 : nullptr,
                                    hrows ? hrows + (size_t)b * N : nullptr, K);
            f[b] = c->eval(x + (size_t)b * n, g + (size_t)b * n);
            delete c;
        }
    });
    return 0;
}

int orc_optimize_batch(const orc_params *p, int B, int N, const double *head, const double *tail,
                       const double *hpolys, const int *hrows, int K, double *x, double *f, int *status,
                       int *iters, int *evals, double *coeffs, double *T, int nthreads) {
    const int S = p->S, n = N + 3 * (N - 1);
    if (S != 3 && S != 4) return -1;
    const orc::LbfgsParams lp = lbfgs_of(*p);
    run_threads(B, nthreads, [&](int lo, int hi) {
        for (int b = lo; b < hi; ++b) {
            CostBase *c = make_cost(*p, N, head + (size_t)b * 3 * S, tail + (size_t)b * 3 * S,
                                    hpolys ? hpolys + (size_t)b * N * K * 4 : nullptr,
                                    hrows ? hrows + (size_t)b * N : nullptr, K);
            orc::LbfgsTrace tr;
            double fb = 0.0;
            const int ret = orc::lbfgs_optimize(n, x + (size_t)b * n, fb, cost_thunk, c, lp, &tr);
            if (f) f[b] = fb;
            if (status) status[b] = ret;
            if (iters) iters[b] = tr.iterations;
            if (evals) evals[b] = tr.evaluations;
            if (coeffs || T)
                c->flat(x + (size_t)b * n, coeffs ? coeffs + (size_t)b * N * 3 * 2 * S : nullptr,
                        T ? T + (size_t)b * N : nullptr);
            delete c;
        }
    });
    return 0;
}

// Same as orc_optimize_batch but the L-BFGS driver is passed in: either orc_lbfgs_optimize (the
// restatement) or ref_lbfgs_optimize from oracle/_ref (the reference's own lbfgs.hpp, verbatim).
typedef int (*lbfgs_driver_fn)(int n, double *x, double *f, orc::eval_fn eval, void *inst, const orc_params *p,
                               int *iters, int *evals);
int orc_optimize_batch_with(lbfgs_driver_fn driver, const orc_params *p, int B, int N, const double *head,
                            const double *tail, const double *hpolys, const int *hrows, int K, double *x, double *f,
                            int *status, int *iters, int *evals, double *coeffs, double *T, int nthreads) {
    const int S = p->S, n = N + 3 * (N - 1);
    if ((S != 3 && S != 4) || !driver) return -1;
    run_threads(B, nthreads, [&](int lo, int hi) {
        for (int b = lo; b < hi; ++b) {
            CostBase *c = make_cost(*p, N, head + (size_t)b * 3 * S, tail + (size_t)b * 3 * S,
                                    hpolys ? hpolys + (size_t)b * N * K * 4 : nullptr,
                                    hrows ? hrows + (size_t)b * N : nullptr, K);
            double fb = 0.0;
            int it = 0, ev = 0;
            const int ret = driver(n, x + (size_t)b * n, &fb, cost_thunk, c, p, &it, &ev);
            if (f) f[b] = fb;
            if (status) status[b] = ret;
            if (iters) iters[b] = it;
            if (evals) evals[b] = ev;
            if (coeffs || T)
                c->flat(x + (size_t)b * n, coeffs ? coeffs + (size_t)b * N * 3 * 2 * S : nullptr,
                        T ? T + (size_t)b * N : nullptr);
            delete c;
        }
    });
    return 0;
}

int orc_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

}  // extern "C"

static void run_threads(int B, int nthreads, const std::function<void(int, int)> &body) {
    if (nthreads <= 1 || B < 2) { body(0, B); return; }
    nthreads = std::min(nthreads, B);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        const int lo = static_cast<int>((long long)B * t / nthreads);
        const int hi = static_cast<int>((long long)B * (t + 1) / nthreads);
        th.emplace_back([=, &body] { body(lo, hi); });
    }
    for (auto &t : th) t.join();
}
