// ============================================================================
// oracle/ref_trajectory.cpp -- C wrapper around the REFERENCE's own Piece<D> / Trajectory<D>.
// TEST INFRASTRUCTURE ONLY.  `#include "gcopter/trajectory.hpp"` resolves to the unmodified file under
// /root/reference/src/planner/include (oracle/Makefile), and so does the "gcopter/root_finder.hpp" it includes
// (Sturm root isolation behind getMaxVelRate / checkMaxVelRate ...); <Eigen/Eigen> is oracle/eigen_shim.  Output: oracle/_ref/libref_lbfgs.so.  It pins the OUTPUT CONTRACT of the hot path:
// coefficients in the library's Trajectory order, fed to the reference's emplace_back(dur, cMat), must
// evaluate (getPos/getVel/getAcc/getJer), locate (locatePieceIdx) and cost (getTrajCost) as the
// reference says -- in particular E_MINCO == 2 * getTrajCost(3) (trajectory.hpp:396-420).
// ============================================================================
#include "gcopter/trajectory.hpp"

extern "C" {

// coeffs [N][3][6], k = 0 highest power (the library's `coeffs` / flatten_coffmats layout); T [N]
void *ref_traj5_create(int N, const double *T, const double *coeffs) {
    Trajectory<5> *tr = new Trajectory<5>();
    for (int i = 0; i < N; ++i) {
        Piece<5>::CoefficientMat c;
        for (int a = 0; a < 3; ++a)
            for (int k = 0; k < 6; ++k) c(a, k) = coeffs[(i * 3 + a) * 6 + k];
        tr->emplace_back(T[i], c);                 // trajectory.hpp:505, as learning_planner.hpp:216 does
    }
    return tr;
}
void ref_traj5_destroy(void *p) { delete static_cast<Trajectory<5> *>(p); }
double ref_traj5_cost(void *p, int order) { return static_cast<Trajectory<5> *>(p)->getTrajCost(order); }
double ref_traj5_total_duration(void *p) { return static_cast<Trajectory<5> *>(p)->getTotalDuration(); }
int ref_traj5_pieces(void *p) { return static_cast<Trajectory<5> *>(p)->getPieceNum(); }
void ref_traj5_eval(void *p, double t, double *pos, double *vel, double *acc, double *jer) {
    Trajectory<5> *tr = static_cast<Trajectory<5> *>(p);
    const Eigen::Vector3d P = tr->getPos(t), V = tr->getVel(t), A = tr->getAcc(t), J = tr->getJer(t);
    for (int a = 0; a < 3; ++a) { pos[a] = P(a); vel[a] = V(a); acc[a] = A(a); jer[a] = J(a); }
}
void ref_traj5_positions(void *p, double *out /* [N+1][3] */) {
    Trajectory<5> *tr = static_cast<Trajectory<5> *>(p);
    const Eigen::Matrix3Xd P = tr->getPositions();
    for (int i = 0; i < P.cols(); ++i)
        for (int a = 0; a < 3; ++a) out[3 * i + a] = P(a, i);
}
int ref_traj5_locate(void *p, double *t_inout) { return static_cast<Trajectory<5> *>(p)->locatePieceIdx(*t_inout); }
// max-rate members (trajectory.hpp:598-646; per piece :177-313) with the reference's own gcopter/root_finder.hpp
double ref_traj5_max_vel_rate(void *p) { return static_cast<Trajectory<5> *>(p)->getMaxVelRate(); }
double ref_traj5_max_acc_rate(void *p) { return static_cast<Trajectory<5> *>(p)->getMaxAccRate(); }
int ref_traj5_check_max_vel_rate(void *p, double v) { return static_cast<Trajectory<5> *>(p)->checkMaxVelRate(v) ? 1 : 0; }
int ref_traj5_check_max_acc_rate(void *p, double a) { return static_cast<Trajectory<5> *>(p)->checkMaxAccRate(a) ? 1 : 0; }

}  // extern "C"
