// oracle/ref_stubs/gcopter/root_finder.hpp -- TEST INFRASTRUCTURE ONLY.
// Placed BEFORE the reference include directory on the compiler's -I list so that the reference's
// gcopter/trajectory.hpp (compiled verbatim from /root/reference by oracle/Makefile) finds this file instead
// of the reference's 1100-line polynomial root finder, which needs the real Eigen.  Only the names that
// trajectory.hpp mentions are declared; the members that use them (getMaxVelRate, checkMaxAccRate, ...)
// are never instantiated by oracle/ref_trajectory.cpp, which exercises exactly the hot-path contract:
// Piece<D>::getPos/getVel/getAcc/getJer, Trajectory<D>::emplace_back/getTrajCost/getPositions/
// getDurations/getTotalDuration/locatePieceIdx.
#pragma once
#include <Eigen/Eigen>

#include <set>

namespace RootFinder {
template <class V> inline Eigen::VectorXd polySqr(const V &) { return Eigen::VectorXd(); }
template <class V> inline double polyVal(const V &, double) { return 0.0; }
template <class V> inline std::set<double> solvePolynomial(const V &, double, double, double) { return std::set<double>(); }
template <class V> inline int countRoots(const V &, double, double) { return 0; }
}  // namespace RootFinder
