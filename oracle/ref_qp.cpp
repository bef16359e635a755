// oracle/ref_qp.cpp -- the reference's incumbent trajectory back-end, src/planner/include/planner/qp_solver.hpp,
// compiled VERBATIM from /root/reference against oracle/qp_stubs (ROS / OsqpEigen / Eigen stand-ins that record what
// QPSolver::solve hands to OSQP).  C entry point: build the QP for one problem and return its matrices.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_qp.so, tests/test_oracle_minco.py).
#include <iostream>

#include "planner/qp_solver.hpp"

extern "C" {

// iniPVA, finPVA: 3x3 row-major (rows = axis, columns P,V,A: learning_planning.cpp:147-151).
// hpolys: seg blocks of rows x 4, rows [n, b] (n.p <= b, learning_planner.hpp:293-299).  times: float32 like the planner's.
// Outputs (row-major, caller-sized): hessian n x n, constraints m x n, lower m, upper m.  Returns 0, or -1 when solve() fails.
int ref_qp_build(int seg, int rows_per_poly, int order, int res, double vel_box, double acc_box, const double *iniPVA,
                 const double *finPVA, const double *hpolys, const float *times, int *n_out, int *m_out, double *hessian,
                 double *constraints, double *lower, double *upper) {
    ros::NodeHandle nh;
    nh.values["MaxVelBox"] = vel_box;
    nh.values["MaxAccBox"] = acc_box;
    nh.values["ConstRes"] = res;
    QPConfig cfg(nh);
    QPSolver solver(cfg);
    solver.setOrder(order);
    Eigen::MatrixXd ini(3, 3), fin(3, 3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { ini(i, j) = iniPVA[i * 3 + j]; fin(i, j) = finPVA[i * 3 + j]; }
    std::vector<Eigen::MatrixX4d> polys;
    for (int s = 0; s < seg; ++s) {
        Eigen::MatrixX4d p(rows_per_poly, 4);
        for (int r = 0; r < rows_per_poly; ++r)
            for (int c = 0; c < 4; ++c) p(r, c) = hpolys[((size_t)s * rows_per_poly + r) * 4 + c];
        polys.push_back(p);
    }
    struct Times {
        const float *t;
        float operator()(size_t i) const { return t[i]; }
    } tm{times};
    Eigen::VectorXd sol;
    std::streambuf *old = std::cout.rdbuf(nullptr);          // the reference prints "[QP solver]: solver success"
    const bool ok = solver.solve(ini, fin, polys, tm, sol);
    std::cout.rdbuf(old);
    const OsqpEigen::Capture &c = OsqpEigen::last_capture();
    *n_out = c.n; *m_out = c.m;
    if (hessian) for (size_t i = 0; i < c.hessian.size(); ++i) hessian[i] = c.hessian.data()[i];
    if (constraints) for (size_t i = 0; i < c.constraints.size(); ++i) constraints[i] = c.constraints.data()[i];
    if (lower) for (size_t i = 0; i < c.lower.size(); ++i) lower[i] = c.lower.data()[i];
    if (upper) for (size_t i = 0; i < c.upper.size(); ++i) upper[i] = c.upper.data()[i];
    return ok ? 0 : -1;
}
}
