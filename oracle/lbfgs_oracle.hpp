// ============================================================================
// oracle/lbfgs_oracle.hpp — CPU restatement of the reference L-BFGS driver.
//
// TEST INFRASTRUCTURE ONLY (see oracle/minco_oracle.hpp header).
//
// Restates, on plain double arrays (Eigen is not installed here):
//   lbfgs_parameter_t             gcopter/lbfgs.hpp:15-129
//   return codes                  gcopter/lbfgs.hpp:135-184
//   line_search_lewisoverton      gcopter/lbfgs.hpp:276-384
//   lbfgs_optimize                gcopter/lbfgs.hpp:434-717
// No deliberate deviation from the reference control flow; progress / stepbound
// callbacks (unused on this path, nullptr at every call site) are omitted.
//
// PINNED against the real reference: tests/test_oracle_lbfgs.py runs this file
// and the verbatim reference header (oracle/_ref/libref_lbfgs.so, built by
// oracle/Makefile from /root/reference/.../lbfgs.hpp against oracle/eigen_shim)
// on the same callbacks and requires identical iterates, step counts and codes.
// ============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace orc {

struct LbfgsParams {
    int mem_size = 8;
    double g_epsilon = 1.0e-5;
    int past = 3;
    double delta = 1.0e-6;
    int max_iterations = 0;
    int max_linesearch = 64;
    double min_step = 1.0e-20;
    double max_step = 1.0e+20;
    double f_dec_coeff = 1.0e-4;
    double s_curv_coeff = 0.9;
    double cautious_factor = 1.0e-6;
    double machine_prec = 1.0e-16;
};

enum {
    LBFGS_CONVERGENCE = 0,
    LBFGS_STOP,
    LBFGS_CANCELED,
    LBFGSERR_UNKNOWNERROR = -1024,
    LBFGSERR_INVALID_N,
    LBFGSERR_INVALID_MEMSIZE,
    LBFGSERR_INVALID_GEPSILON,
    LBFGSERR_INVALID_TESTPERIOD,
    LBFGSERR_INVALID_DELTA,
    LBFGSERR_INVALID_MINSTEP,
    LBFGSERR_INVALID_MAXSTEP,
    LBFGSERR_INVALID_FDECCOEFF,
    LBFGSERR_INVALID_SCURVCOEFF,
    LBFGSERR_INVALID_MACHINEPREC,
    LBFGSERR_INVALID_MAXLINESEARCH,
    LBFGSERR_INVALID_FUNCVAL,
    LBFGSERR_MINIMUMSTEP,
    LBFGSERR_MAXIMUMSTEP,
    LBFGSERR_MAXIMUMLINESEARCH,
    LBFGSERR_MAXIMUMITERATION,
    LBFGSERR_WIDTHTOOSMALL,
    LBFGSERR_INVALIDPARAMETERS,
    LBFGSERR_INCREASEGRADIENT,
};

typedef double (*eval_fn)(void *instance, const double *x, double *g, int n);

struct LbfgsTrace {  // what the parity tests compare
    int iterations = 0;   // k at exit
    int evaluations = 0;  // callback invocations
};

namespace detail {
inline double dot(const double *a, const double *b, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
inline double inf_norm(const double *a, int n) {
    double m = 0.0;
    for (int i = 0; i < n; ++i) m = std::max(m, std::fabs(a[i]));
    return m;
}
}  // namespace detail

// Lewis-Overton weak-Wolfe search.  Returns evaluation count (>0) or an error.
inline int line_search_lo(int n, double *x, double &f, double *g, double &stp, const double *s,
                          const double *xp, const double *gp, double stpmin, double stpmax,
                          eval_fn eval, void *inst, const LbfgsParams &pr, int &evals) {
    if (!(stp > 0.0)) return LBFGSERR_INVALIDPARAMETERS;
    const double dginit = detail::dot(gp, s, n);
    if (0.0 < dginit) return LBFGSERR_INCREASEGRADIENT;
    const double finit = f;
    const double dgtest = pr.f_dec_coeff * dginit;
    const double dstest = pr.s_curv_coeff * dginit;
    int count = 0;
    bool bracketed = false, touched = false;
    double lo = 0.0, hi = stpmax;
    for (;;) {
        for (int i = 0; i < n; ++i) x[i] = xp[i] + stp * s[i];
        f = eval(inst, x, g, n);
        ++count; ++evals;
        if (std::isinf(f) || std::isnan(f)) return LBFGSERR_INVALID_FUNCVAL;
        if (f > finit + stp * dgtest) {           // Armijo fails: shrink from above
            hi = stp; bracketed = true;
        } else if (detail::dot(g, s, n) < dstest) {  // weak Wolfe fails: grow from below
            lo = stp;
        } else {
            return count;
        }
        if (pr.max_linesearch <= count) return LBFGSERR_MAXIMUMLINESEARCH;
        if (bracketed && (hi - lo) < pr.machine_prec * hi) return LBFGSERR_WIDTHTOOSMALL;
        stp = bracketed ? 0.5 * (lo + hi) : stp * 2.0;
        if (stp < stpmin) return LBFGSERR_MINIMUMSTEP;
        if (stp > stpmax) {
            if (touched) return LBFGSERR_MAXIMUMSTEP;
            touched = true;
            stp = stpmax;
        }
    }
}

inline int lbfgs_optimize(int n, double *x, double &f, eval_fn eval, void *inst,
                          const LbfgsParams &pr, LbfgsTrace *trace = nullptr) {
    const int m = pr.mem_size;
    if (n <= 0) return LBFGSERR_INVALID_N;
    if (m <= 0) return LBFGSERR_INVALID_MEMSIZE;
    if (pr.g_epsilon < 0.0) return LBFGSERR_INVALID_GEPSILON;
    if (pr.past < 0) return LBFGSERR_INVALID_TESTPERIOD;
    if (pr.delta < 0.0) return LBFGSERR_INVALID_DELTA;
    if (pr.min_step < 0.0) return LBFGSERR_INVALID_MINSTEP;
    if (pr.max_step < pr.min_step) return LBFGSERR_INVALID_MAXSTEP;
    if (!(pr.f_dec_coeff > 0.0 && pr.f_dec_coeff < 1.0)) return LBFGSERR_INVALID_FDECCOEFF;
    if (!(pr.s_curv_coeff < 1.0 && pr.s_curv_coeff > pr.f_dec_coeff)) return LBFGSERR_INVALID_SCURVCOEFF;
    if (!(pr.machine_prec > 0.0)) return LBFGSERR_INVALID_MACHINEPREC;
    if (pr.max_linesearch <= 0) return LBFGSERR_INVALID_MAXLINESEARCH;

    std::vector<double> xp(n), g(n), gp(n), d(n), pf(std::max(1, pr.past));
    std::vector<double> alpha(m, 0.0), ysv(m, 0.0);
    std::vector<double> ms(static_cast<size_t>(n) * m, 0.0), my(static_cast<size_t>(n) * m, 0.0);  // column c at [c*n]
    int evals = 0, ret = 0, k = 0;

    double fx = eval(inst, x, g.data(), n);
    ++evals;
    pf[0] = fx;
    for (int i = 0; i < n; ++i) d[i] = -g[i];

    double gn = detail::inf_norm(g.data(), n), xn = detail::inf_norm(x, n);
    if (gn / std::max(1.0, xn) < pr.g_epsilon) {
        ret = LBFGS_CONVERGENCE;
    } else {
        double step = 1.0 / std::sqrt(detail::dot(d.data(), d.data(), n));
        k = 1;
        int end = 0, bound = 0;
        for (;;) {
            xp.assign(x, x + n);
            gp = g;
            const int ls = line_search_lo(n, x, fx, g.data(), step, d.data(), xp.data(), gp.data(),
                                          pr.min_step, pr.max_step, eval, inst, pr, evals);
            if (ls < 0) {
                std::copy(xp.begin(), xp.end(), x);
                g = gp;
                ret = ls;
                break;
            }
            gn = detail::inf_norm(g.data(), n);
            xn = detail::inf_norm(x, n);
            if (gn / std::max(1.0, xn) < pr.g_epsilon) { ret = LBFGS_CONVERGENCE; break; }
            if (0 < pr.past) {
                if (pr.past <= k) {
                    const double rate = std::fabs(pf[k % pr.past] - fx) / std::max(1.0, std::fabs(fx));
                    if (rate < pr.delta) { ret = LBFGS_STOP; break; }
                }
                pf[k % pr.past] = fx;
            }
            if (pr.max_iterations != 0 && pr.max_iterations <= k) { ret = LBFGSERR_MAXIMUMITERATION; break; }
            ++k;
            double *sc = &ms[static_cast<size_t>(end) * n], *yc = &my[static_cast<size_t>(end) * n];
            for (int i = 0; i < n; ++i) { sc[i] = x[i] - xp[i]; yc[i] = g[i] - gp[i]; }
            const double ys = detail::dot(yc, sc, n);
            const double yy = detail::dot(yc, yc, n);
            ysv[end] = ys;
            for (int i = 0; i < n; ++i) d[i] = -g[i];
            const double cau = detail::dot(sc, sc, n) * std::sqrt(detail::dot(gp.data(), gp.data(), n)) * pr.cautious_factor;
            if (ys > cau) {
                ++bound;
                bound = m < bound ? m : bound;
                end = (end + 1) % m;
                int j = end;
                for (int i = 0; i < bound; ++i) {
                    j = (j + m - 1) % m;
                    alpha[j] = detail::dot(&ms[static_cast<size_t>(j) * n], d.data(), n) / ysv[j];
                    const double na = -alpha[j];
                    for (int u = 0; u < n; ++u) d[u] += na * my[static_cast<size_t>(j) * n + u];
                }
                const double sc0 = ys / yy;
                for (int u = 0; u < n; ++u) d[u] *= sc0;
                for (int i = 0; i < bound; ++i) {
                    const double beta = detail::dot(&my[static_cast<size_t>(j) * n], d.data(), n) / ysv[j];
                    const double cf = alpha[j] - beta;
                    for (int u = 0; u < n; ++u) d[u] += cf * ms[static_cast<size_t>(j) * n + u];
                    j = (j + 1) % m;
                }
            }
            step = 1.0;
        }
    }
    f = fx;
    if (trace) { trace->iterations = k; trace->evaluations = evals; }
    return ret;
}

}  // namespace orc
