// ============================================================================
// include/mincob/sfc_optimizer.hpp -- C++ host side of the trajectory back-end slot.
//
//   mincob::solve(iniPVA, finPVA, hPolys, times, flatten_coffmats)
//       same argument list, return convention (bool) and output layout (idx = i*3*d + j*d + k,
//       k = 0 highest power) as the call it replaces:
//           qp_solver.solve(iniPVA, finPVA, hPolys, times, flatten_coffmats)
//           src/planner/include/planner/learning_planner.hpp:196   (consumer :201-233)
//       Two time modes (last argument):
//         TimeMode::Fixed (default)  the literal replacement: the network's durations are DATA, exactly as in the
//             incumbent QP (qp_solver.hpp:119-360 keeps `times` fixed); only the waypoints are optimised
//             (MINCOB_FLAG_FREEZE_TIMES) and `times` is not written.
//         TimeMode::Optimize         `times` is in/out: the net's allocation is the WARM START of the spatial-temporal
//             optimisation (upstream GCOPTER behaviour) and the optimised durations are written back, so that
//             `jerk_traj.emplace_back(times(i), coffMat)` at :216 builds the optimised trajectory.
//
//   mincob::PolytopeSFC   setup(...) / optimize(...) / static costFunctional(void*, x, g)
//       the upstream GCOPTER_PolytopeSFC shape: `costFunctional` has the
//       lbfgs_evaluate_t signature of gcopter/lbfgs.hpp:200-202, so the reference's own
//           lbfgs::lbfgs_optimize(x, f, &PolytopeSFC::costFunctional, nullptr, nullptr, &sfc, params)
//       runs unchanged with the cost evaluated by the sm_100a kernel (one launch per call);
//       `optimize` runs the whole L-BFGS loop on the device instead (one launch in total).
//
//   mincob::BatchOptimizer   B problems at once (host pointers in the layouts of include/mincob.h).
//
// Half-plane sign: the planner hands rows [n, b] with n.p <= b (learning_planner.hpp:293-299); they are passed on
// unchanged with MINCOB_FLAG_PLANNER_ROWS set (GCOPTER's own n.p + d <= 0 form, geo_utils.hpp:41-42, is the
// library's default for callers that come from sfc_gen / firi directly).
// ============================================================================
#pragma once
#include <Eigen/Eigen>

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

#include "minco.hpp"

namespace mincob {

inline double forwardT(double tau) { return tau > 0.0 ? (0.5 * tau + 1.0) * tau + 1.0 : 1.0 / ((0.5 * tau - 1.0) * tau + 1.0); }
inline double backwardT(double T) { return T > 1.0 ? std::sqrt(2.0 * T - 1.0) - 1.0 : 1.0 - std::sqrt(2.0 / T - 1.0); }

// A point inside consecutive polytopes i-1 and i (they overlap along a corridor) by cyclic projection
// onto the violated half-planes, started from `p`.  Only an initial guess for the optimiser.
template <class Polys>
inline void projectIntoOverlap(const Polys &hPolys, int i, double margin, double p[3]) {
    for (int sweep = 0; sweep < 64; ++sweep) {
        bool moved = false;
        for (int which = i - 1; which <= i; ++which) {
            const auto &H = hPolys[which];
            for (int r = 0; r < (int)H.rows(); ++r) {
                const double nn = H(r, 0) * H(r, 0) + H(r, 1) * H(r, 1) + H(r, 2) * H(r, 2);
                if (nn == 0.0) continue;
                const double v = H(r, 0) * p[0] + H(r, 1) * p[1] + H(r, 2) * p[2] - H(r, 3) + margin * std::sqrt(nn);
                if (v > 0.0) {
                    for (int a = 0; a < 3; ++a) p[a] -= v / nn * H(r, a);
                    moved = true;
                }
            }
        }
        if (!moved) break;
    }
}

class PolytopeSFC {
public:
    ~PolytopeSFC() { if (h) mincob_destroy(h); }

    // iniPVA / finPVA: 3 x 3 (rows axis, columns P,V,A); hPolys[i]: rows [n, b], n.p <= b, one polytope per
    // piece (qp_solver.hpp:126); times0: initial durations (> 0); inPs0: optional 3 x (N-1) initial waypoints.
    template <class MatPVA, class Polys, class Times>
    inline bool setup(const MatPVA &iniPVA, const MatPVA &finPVA, const Polys &hPolys, const Times &times0,
                      const Eigen::Matrix3Xd *inPs0 = nullptr, const mincob_params *params = nullptr, int device = 0) {
        N = (int)hPolys.size();
        if (N < 1 || N > MINCOB_MAX_PIECES) return false;
        if (params) prm = *params; else mincob_default_params(&prm, 3);
        prm.flags |= MINCOB_FLAG_PLANNER_ROWS;                  // hPolys rows are [n, b], n.p <= b
        const int S = prm.S;
        if (!h && mincob_create(&h, &prm, device) != 0) return false;
        if (mincob_set_params(h, &prm) != 0) return false;
        K = 0;
        for (int i = 0; i < N; ++i) K = std::max(K, (int)hPolys[i].rows());
        head.assign(3 * S, 0.0); tail.assign(3 * S, 0.0);
        for (int d = 0; d < std::min(S, (int)iniPVA.cols()); ++d)
            for (int a = 0; a < 3; ++a) { head[3 * d + a] = iniPVA(a, d); tail[3 * d + a] = finPVA(a, d); }
        planes.assign((size_t)N * std::max(K, 1) * 4, 0.0);
        rows.assign(N, 0);
        for (int i = 0; i < N; ++i) {
            rows[i] = (int)hPolys[i].rows();
            for (int r = 0; r < rows[i]; ++r) {
                double *o = &planes[((size_t)i * K + r) * 4];
                o[0] = hPolys[i](r, 0); o[1] = hPolys[i](r, 1); o[2] = hPolys[i](r, 2); o[3] = hPolys[i](r, 3);
            }
        }
        n = N + 3 * (N - 1);
        x.assign(n, 0.0);
        for (int i = 0; i < N; ++i) {
            const double t = times0(i);
            if (!(t > 0.0)) return false;                       // learning_planner.hpp:181-189
            x[i] = backwardT(t);
        }
        for (int i = 1; i < N; ++i) {
            double p[3];
            for (int a = 0; a < 3; ++a)
                p[a] = inPs0 ? (*inPs0)(a, i - 1) : head[a] + (tail[a] - head[a]) * (double)i / N;
            if (!inPs0) projectIntoOverlap(hPolys, i, 0.05, p);
            for (int a = 0; a < 3; ++a) x[N + 3 * (i - 1) + a] = p[a];
        }
        return mincob_set_problems(h, 1, N, K, head.data(), tail.data(), K > 0 ? planes.data() : nullptr,
                                   K > 0 ? rows.data() : nullptr) == 0;
    }

    // lbfgs_optimize on the device.  Returns the final cost; status gets the lbfgs.hpp return code.
    inline double optimize(int *status = nullptr) {
        const int S = prm.S;
        coeffs.assign((size_t)N * 3 * 2 * S, 0.0);
        durations.assign(N, 0.0);
        double f = 0.0;
        int32_t st = 0, it = 0, ev = 0;
        const int rc = mincob_optimize(h, x.data(), &f, &st, &it, &ev, coeffs.data(), durations.data());
        lastStatus = rc != 0 ? -1024 : st;
        iterations = it; evaluations = ev;
        if (status) *status = lastStatus;
        return f;
    }

    // lbfgs_evaluate_t (gcopter/lbfgs.hpp:200-202): instance is a PolytopeSFC*
    static inline double costFunctional(void *ptr, const Eigen::VectorXd &xv, Eigen::VectorXd &g) {
        PolytopeSFC &o = *static_cast<PolytopeSFC *>(ptr);
        std::vector<double> xin(o.n), gout(o.n);
        for (int i = 0; i < o.n; ++i) xin[i] = xv(i);
        double f = 0.0;
        check(o.h, mincob_evaluate(o.h, xin.data(), &f, gout.data()), "costFunctional");
        for (int i = 0; i < o.n; ++i) g(i) = gout[i];
        return f;
    }

    int N = 0, K = 0, n = 0, lastStatus = 0, iterations = 0, evaluations = 0;
    mincob_params prm;
    std::vector<double> x, coeffs, durations;   // x = [tau; q]; coeffs [N][3][2S] Trajectory order

private:
    mincob_handle h = nullptr;
    std::vector<double> head, tail, planes;
    std::vector<int32_t> rows;
};

enum class TimeMode { Fixed, Optimize };

// Drop-in for qp_solver.solve at learning_planner.hpp:196 (see the header comment).
template <class MatPVA, class Polys, class Times>
inline bool solve(const MatPVA &iniPVA, const MatPVA &finPVA, const Polys &hPolys, Times &times,
                  Eigen::VectorXd &flat_coeffs, const Eigen::Matrix3Xd *inPs0 = nullptr,
                  const mincob_params *params = nullptr, TimeMode mode = TimeMode::Fixed) {
    static PolytopeSFC sfc;   // LearningPlanner is single-threaded and not re-entrant
    mincob_params p;
    if (params) p = *params; else mincob_default_params(&p, 3);
    if (mode == TimeMode::Fixed) p.flags |= MINCOB_FLAG_FREEZE_TIMES; else p.flags &= ~MINCOB_FLAG_FREEZE_TIMES;
    if (!sfc.setup(iniPVA, finPVA, hPolys, times, inPs0, &p)) return false;
    int status = 0;
    sfc.optimize(&status);
    if (status < 0 && status != -1008 /* LBFGSERR_MAXIMUMITERATION: best iterate is still usable */) return false;
    flat_coeffs.resize((int)sfc.coeffs.size());
    for (size_t i = 0; i < sfc.coeffs.size(); ++i) flat_coeffs((int)i) = sfc.coeffs[i];
    if (mode == TimeMode::Optimize) {
        typedef typename std::decay<decltype(times(0))>::type TimeScalar;   // float for the net's VectorXf
        for (int i = 0; i < sfc.N; ++i) times(i) = (TimeScalar)sfc.durations[i];
    }
    return true;
}

// B problems at once; thin RAII over the C-ABI for C++ callers that own host arrays in its layouts.
class BatchOptimizer {
public:
    explicit BatchOptimizer(const mincob_params &p, int device = 0) {
        const int rc = mincob_create(&h, &p, device);
        if (rc != 0) throw std::runtime_error(std::string("mincob_create: ") + mincob_strerror(rc));
    }
    ~BatchOptimizer() { if (h) mincob_destroy(h); }
    BatchOptimizer(const BatchOptimizer &) = delete;
    BatchOptimizer &operator=(const BatchOptimizer &) = delete;
    void setProblems(int B, int N, int K, const double *head, const double *tail, const double *hpolys, const int32_t *hrows) {
        check(h, mincob_set_problems(h, B, N, K, head, tail, hpolys, hrows), "setProblems");
    }
    // overlapped upload for large batches: page-locked arrays (mincob_host_alloc / mincob_host_register) that stay untouched
    // until the following optimize() has returned; the kernel starts while the batch is still arriving
    void setProblemsAsync(int B, int N, int K, const double *head, const double *tail, const double *hpolys, const int32_t *hrows) {
        check(h, mincob_set_problems_async(h, B, N, K, head, tail, hpolys, hrows), "setProblemsAsync");
    }
    void evaluate(const double *x, double *f, double *g) { check(h, mincob_evaluate(h, x, f, g), "evaluate"); }
    void optimize(double *x, double *f, int32_t *status, int32_t *iters, int32_t *evals, double *coeffs, double *T) {
        check(h, mincob_optimize(h, x, f, status, iters, evals, coeffs, T), "optimize");
    }
    // what Trajectory<D>::getMaxVelRate / getMaxAccRate (gcopter/trajectory.hpp:598-622) return, per trajectory, plus the
    // same for jerk: rates [B][3]; checkMaxVelRate(v) / checkMaxAccRate(a) are rates[b][0] < v / rates[b][1] < a
    void maxRates(const double *coeffs, const double *T, double *rates) {
        check(h, mincob_max_rates(h, coeffs, T, rates), "maxRates");
    }
    // sampled max |v|, |a|, |j| and the largest corridor residual (<= 0: inside): report [B][4]
    void checkFeasibility(const double *coeffs, const double *T, int samples, double *report) {
        check(h, mincob_check_feasibility(h, coeffs, T, samples, report), "checkFeasibility");
    }
    mincob_handle handle() const { return h; }

private:
    mincob_handle h = nullptr;
};

}  // namespace mincob
