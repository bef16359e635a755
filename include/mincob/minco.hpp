// ============================================================================
// include/mincob/minco.hpp -- C++ host side of the MINCO classes named by BASELINE.json:
//   minco::MINCO_S3NU / minco::MINCO_S4NU with the upstream GCOPTER method names and argument
//   types (setConditions, setParameters, getTrajectory, getEnergy, getCoeffs,
//   getEnergyPartialGradByCoeffs, getEnergyPartialGradByTimes, propogateGrad [sic]; restated in
//   SURVEY.md Appendix A -- that header is NOT vendored in the reference).
// Header-only, like every gcopter/*.hpp of the reference; the arithmetic runs in the sm_100a
// kernels of libmincob.so through the C-ABI of include/mincob.h (B = 1 here; the batched entry
// points are what a throughput caller uses, see sfc_optimizer.hpp).  There is no CPU path: a
// failing C-ABI call throws std::runtime_error with mincob_last_error().
//
// Eigen: only rows()/cols()/size()/resize()/operator() are used, so the same source compiles
// against real Eigen (ROS box) and against the stand-in used by tests/cpp in this repository.
// Layout notes (parity on indexing): headState/tailState are 3 x S with columns P,V,A[,J]
// (learning_planning.cpp:147-151); inPs is 3 x (N-1); getCoeffs() is (2S*N) x 3 with row
// 2S*i + k = c_k of piece i, ASCENDING powers; getTrajectory() emits 3 x 2S blocks in DESCENDING
// powers, the Piece<D> order of gcopter/trajectory.hpp:79-83.
// ============================================================================
#pragma once
#include <Eigen/Eigen>

#include <stdexcept>
#include <string>
#include <vector>

#include "../mincob.h"

namespace mincob {

// One lazily created handle per order S and process: the planner is single-threaded
// (learning_planning.cpp:314-320) and must not have to manage CUDA contexts or streams.
inline mincob_handle shared_handle(int S, int device = 0) {
    static mincob_handle h[2] = {nullptr, nullptr};
    mincob_handle &slot = h[S == 3 ? 0 : 1];
    if (!slot) {
        mincob_params p;
        if (mincob_default_params(&p, S) != 0) throw std::runtime_error("mincob: S must be 3 or 4");
        const int rc = mincob_create(&slot, &p, device);
        if (rc != 0) throw std::runtime_error(std::string("mincob_create: ") + mincob_strerror(rc));
    }
    return slot;
}
inline void check(mincob_handle h, int rc, const char *what) {
    if (rc != 0)
        throw std::runtime_error(std::string(what) + ": " + mincob_strerror(rc) + " (" + mincob_last_error(h) + ")");
}

}  // namespace mincob

namespace minco {

template <int S>
class MINCO_SNU {
public:
    typedef Eigen::Matrix<double, 3, S> BoundaryMat;   // Matrix3d for S = 3 (headPVA), 3x4 for S = 4

    MINCO_SNU() = default;

    inline void setConditions(const BoundaryMat &headState, const BoundaryMat &tailState, const int &pieceNum) {
        N = pieceNum;
        head.resize(3 * S); tail.resize(3 * S);
        for (int d = 0; d < S; ++d)
            for (int a = 0; a < 3; ++a) { head[3 * d + a] = headState(a, d); tail[3 * d + a] = tailState(a, d); }
        b.resize(2 * S * N, 3);
        gdC.resize(2 * S * N, 3);
        flat.assign((size_t)N * 3 * 2 * S, 0.0);
        T1.resize(N);
        gdT.resize(N);
    }

    // banded solve for the coefficients; energy and its partials come out of the same launch
    inline void setParameters(const Eigen::Matrix3Xd &inPs, const Eigen::VectorXd &ts) {
        mincob_handle h = mincob::shared_handle(S);
        q.resize((size_t)3 * (N > 1 ? N - 1 : 1));
        for (int i = 0; i < N - 1; ++i)
            for (int a = 0; a < 3; ++a) q[3 * i + a] = inPs(a, i);
        std::vector<double> t(N), c((size_t)2 * S * N * 3), gc((size_t)2 * S * N * 3), gt(N);
        for (int i = 0; i < N; ++i) { t[i] = ts(i); T1(i) = ts(i); }
        mincob::check(h, mincob_minco_forward(h, 1, N, head.data(), tail.data(), q.data(), t.data(), c.data(), &energy,
                                              gc.data(), gt.data(), flat.data()),
                      "MINCO::setParameters");
        for (int r = 0; r < 2 * S * N; ++r)
            for (int a = 0; a < 3; ++a) { b(r, a) = c[(size_t)r * 3 + a]; gdC(r, a) = gc[(size_t)r * 3 + a]; }
        for (int i = 0; i < N; ++i) gdT(i) = gt[i];
        times = t;
    }

    // Traj is gcopter's Trajectory<2S-1>: emplace_back(duration, 3 x 2S coefficient matrix, descending powers)
    template <class Traj>
    inline void getTrajectory(Traj &traj) const {
        traj.clear();
        traj.reserve(N);
        Eigen::Matrix<double, 3, 2 * S> cMat;
        for (int i = 0; i < N; ++i) {
            for (int a = 0; a < 3; ++a)
                for (int k = 0; k < 2 * S; ++k) cMat(a, k) = flat[((size_t)i * 3 + a) * 2 * S + k];
            traj.emplace_back(T1(i), cMat);
        }
    }

    inline void getEnergy(double &e) const { e = energy; }
    inline const Eigen::MatrixX3d &getCoeffs(void) const { return b; }
    inline void getEnergyPartialGradByCoeffs(Eigen::MatrixX3d &out) const { out = gdC; }
    inline void getEnergyPartialGradByTimes(Eigen::VectorXd &out) const { out = gdT; }

    // adjoint banded solve: partial dJ/dc, dJ/dT  ->  total dJ/dq, dJ/dT
    inline void propogateGrad(const Eigen::MatrixX3d &partialGradByCoeffs, const Eigen::VectorXd &partialGradByTimes,
                              Eigen::Matrix3Xd &gradByPoints, Eigen::VectorXd &gradByTimes) {
        mincob_handle h = mincob::shared_handle(S);
        std::vector<double> gc((size_t)2 * S * N * 3), gt(N), gq((size_t)3 * (N > 1 ? N - 1 : 1)), gT(N);
        for (int r = 0; r < 2 * S * N; ++r)
            for (int a = 0; a < 3; ++a) gc[(size_t)r * 3 + a] = partialGradByCoeffs(r, a);
        for (int i = 0; i < N; ++i) gt[i] = partialGradByTimes(i);
        mincob::check(h, mincob_minco_propagate(h, 1, N, head.data(), tail.data(), q.data(), times.data(), gc.data(),
                                                gt.data(), gq.data(), gT.data()),
                      "MINCO::propogateGrad");
        gradByPoints.resize(3, N - 1);
        gradByTimes.resize(N);
        for (int i = 0; i < N - 1; ++i)
            for (int a = 0; a < 3; ++a) gradByPoints(a, i) = gq[3 * i + a];
        for (int i = 0; i < N; ++i) gradByTimes(i) = gT[i];
    }

private:
    int N = 0;
    std::vector<double> head, tail, q, times, flat;
    Eigen::MatrixX3d b, gdC;
    Eigen::VectorXd T1, gdT;
    double energy = 0.0;
};

typedef MINCO_SNU<3> MINCO_S3NU;   // quintic pieces, jerk energy: Trajectory<5> (reference default, planner.yaml:23)
typedef MINCO_SNU<4> MINCO_S4NU;   // septic pieces, snap energy: Trajectory<7>

}  // namespace minco
