// ============================================================================
// oracle/ref_firi.cpp -- C wrapper around the REFERENCE's own smoothedL1.
// TEST INFRASTRUCTURE ONLY.  firi.hpp as a whole needs the real Eigen, sdlp and geo_utils, but the
// function itself (src/planner/include/gcopter/firi.hpp:60-84) is plain C++: oracle/Makefile cuts
// exactly those lines out of the reference file where it lies into a temporary include (checked to
// start with the function's signature, deleted after the compile, never committed) and this file
// compiles them unmodified into oracle/_ref/libref_lbfgs.so.
// ============================================================================
namespace firi_ref {
#include "firi_smoothedL1.inc"
}

extern "C" int ref_smoothed_l1(double mu, double x, double *f, double *df) {
    return firi_ref::smoothedL1(mu, x, *f, *df) ? 1 : 0;
}
