// tests/cpp/test_shim.cpp -- drives the C++ host headers (include/mincob/*.hpp) the way the planner
// would, on one problem read from a text file, and prints what it got as one JSON object.
// tests/test_cpp_shim.py compares that with the CPU oracle.  Built by __graft_entry__.build():
//   tests/cpp/_build/test_shim            (our headers + the Eigen stand-in)
//   oracle/_ref/test_shim_reflbfgs        (-DHAVE_REF_LBFGS: additionally the reference's own
//                                          gcopter/lbfgs.hpp, verbatim, driving our GPU costFunctional)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#ifdef HAVE_REF_LBFGS
#include "gcopter/lbfgs.hpp"
#endif
#include "mincob/sfc_optimizer.hpp"

// what Trajectory<5> looks like to getTrajectory (gcopter/trajectory.hpp:505: emplace_back(dur, cMat))
struct FakeTrajectory {
    std::vector<double> dur;
    std::vector<Eigen::Matrix<double, 3, 6>> mats;
    void clear() { dur.clear(); mats.clear(); }
    void reserve(int) {}
    void emplace_back(const double &d, const Eigen::Matrix<double, 3, 6> &m) { dur.push_back(d); mats.push_back(m); }
};

static void dump(const char *name, const std::vector<double> &v, bool comma = true) {
    std::printf("\"%s\": [", name);
    for (size_t i = 0; i < v.size(); ++i) std::printf("%s%.17g", i ? ", " : "", v[i]);
    std::printf("]%s\n", comma ? "," : "");
}

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: test_shim problem.txt\n"); return 2; }
    std::ifstream in(argv[1]);
    int N, K;
    in >> N >> K;
    Eigen::Matrix3d ini, fin;
    for (int d = 0; d < 3; ++d) for (int a = 0; a < 3; ++a) in >> ini(a, d);
    for (int d = 0; d < 3; ++d) for (int a = 0; a < 3; ++a) in >> fin(a, d);
    Eigen::VectorXf times(5 > N ? 5 : N);            // the planner's Map<VectorXf>(.., 5), learning_planner.hpp:179
    Eigen::VectorXd ts(N);
    for (int i = 0; i < N; ++i) { double t; in >> t; times(i) = (float)t; ts(i) = (double)times(i); }
    Eigen::Matrix3Xd q0(3, N - 1);
    for (int i = 0; i < N - 1; ++i) for (int a = 0; a < 3; ++a) in >> q0(a, i);
    std::vector<Eigen::MatrixX4d> hPolys(N);
    for (int i = 0; i < N; ++i) {
        int rows; in >> rows;
        hPolys[i].resize(rows, 4);
        for (int r = 0; r < rows; ++r) for (int c = 0; c < 4; ++c) in >> hPolys[i](r, c);   // planner form [n, b]
    }
    if (!in) { std::fprintf(stderr, "bad problem file\n"); return 2; }
    try {
        std::printf("{\n");
        // 1. MINCO_S3NU, upstream call sequence
        minco::MINCO_S3NU minco;
        minco.setConditions(ini, fin, N);
        minco.setParameters(q0, ts);
        double E; minco.getEnergy(E);
        Eigen::MatrixX3d gdC; Eigen::VectorXd gdT;
        minco.getEnergyPartialGradByCoeffs(gdC);
        minco.getEnergyPartialGradByTimes(gdT);
        Eigen::Matrix3Xd gq; Eigen::VectorXd gT;
        minco.propogateGrad(gdC, gdT, gq, gT);
        FakeTrajectory traj; minco.getTrajectory(traj);
        std::vector<double> v;
        std::printf("\"energy\": %.17g,\n", E);
        const Eigen::MatrixX3d &b = minco.getCoeffs();
        v.clear(); for (int r = 0; r < 6 * N; ++r) for (int a = 0; a < 3; ++a) v.push_back(b(r, a)); dump("coeffs_asc", v);
        v.clear(); for (int i = 0; i < N - 1; ++i) for (int a = 0; a < 3; ++a) v.push_back(gq(a, i)); dump("gradByPoints", v);
        v.clear(); for (int i = 0; i < N; ++i) v.push_back(gT(i)); dump("gradByTimes", v);
        v.clear(); for (int i = 0; i < N; ++i) for (int a = 0; a < 3; ++a) for (int k = 0; k < 6; ++k) v.push_back(traj.mats[i](a, k)); dump("traj_desc", v);
        // 2. costFunctional at the start point
        mincob::PolytopeSFC sfc;
        if (!sfc.setup(ini, fin, hPolys, times, &q0)) { std::fprintf(stderr, "setup failed\n"); return 1; }
        Eigen::VectorXd x(sfc.n), g(sfc.n);
        for (int i = 0; i < sfc.n; ++i) x(i) = sfc.x[i];
        const double f0 = mincob::PolytopeSFC::costFunctional(&sfc, x, g);
        std::printf("\"f0\": %.17g,\n", f0);
        v.assign(sfc.x.begin(), sfc.x.end()); dump("x0", v);
        v.clear(); for (int i = 0; i < sfc.n; ++i) v.push_back(g(i)); dump("g0", v);
#ifdef HAVE_REF_LBFGS
        // 3. the reference's own lbfgs_optimize around the GPU costFunctional (upstream optimize() settings)
        {
            lbfgs::lbfgs_parameter_t lp;
            lp.mem_size = sfc.prm.mem_size; lp.past = sfc.prm.past; lp.g_epsilon = sfc.prm.g_epsilon;
            lp.min_step = sfc.prm.min_step; lp.delta = sfc.prm.delta; lp.max_iterations = sfc.prm.max_iterations;
            Eigen::VectorXd xh = x;
            double fh = 0.0;
            const int ret = lbfgs::lbfgs_optimize(xh, fh, &mincob::PolytopeSFC::costFunctional, nullptr, nullptr, &sfc, lp);
            std::printf("\"host_lbfgs_ret\": %d, \"host_lbfgs_f\": %.17g,\n", ret, fh);
            v.clear(); for (int i = 0; i < sfc.n; ++i) v.push_back(xh(i)); dump("host_lbfgs_x", v);
        }
#endif
        // 4. the same loop on the device
        int status = 0;
        const double f = sfc.optimize(&status);
        std::printf("\"dev_f\": %.17g, \"dev_status\": %d, \"dev_iters\": %d, \"dev_evals\": %d,\n", f, status, sfc.iterations, sfc.evaluations);
        dump("dev_x", sfc.x); dump("dev_coeffs", sfc.coeffs); dump("dev_T", sfc.durations);
        // 5. the one-line replacement of qp_solver.solve (learning_planner.hpp:196); q0 by projection
        Eigen::VectorXd flat;
        {   // 5a. literal replacement: durations fixed, `times` untouched
            Eigen::VectorXd flat_fixed;
            const bool okf = mincob::solve(ini, fin, hPolys, times, flat_fixed);
            std::printf("\"solvefix_ok\": %s,\n", okf ? "true" : "false");
            v.clear(); for (int i = 0; i < N; ++i) v.push_back((double)times(i)); dump("solvefix_times", v);
            v.clear(); for (int i = 0; i < flat_fixed.size(); ++i) v.push_back(flat_fixed(i)); dump("solvefix_flat", v);
        }
        // 5b. durations refined as well (times in/out)
        const bool ok = mincob::solve(ini, fin, hPolys, times, flat, nullptr, nullptr, mincob::TimeMode::Optimize);
        std::printf("\"solve_ok\": %s,\n", ok ? "true" : "false");
        v.clear(); for (int i = 0; i < N; ++i) v.push_back((double)times(i)); dump("solve_times", v);
        v.clear(); for (int i = 0; i < flat.size(); ++i) v.push_back(flat(i)); dump("solve_flat", v, false);
        std::printf("}\n");
    } catch (const std::exception &e) {
        std::fprintf(stderr, "mincob error: %s\n", e.what());
        return 3;
    }
    return 0;
}
