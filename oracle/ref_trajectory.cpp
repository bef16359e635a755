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

// One set of C entry points per trajectory degree: ref_traj5_* wraps Trajectory<5> (MINCO_S3NU, quintic pieces),
// ref_traj7_* wraps Trajectory<7> (MINCO_S4NU, septic pieces).  coeffs [N][3][D+1], k = 0 highest power (the
// library's `coeffs` / flatten_coffmats layout); T [N].
#define REF_TRAJ_API(DEG)                                                                                          \
    extern "C" void *ref_traj##DEG##_create(int N, const double *T, const double *coeffs) {                        \
        Trajectory<DEG> *tr = new Trajectory<DEG>();                                                               \
        for (int i = 0; i < N; ++i) {                                                                              \
            Piece<DEG>::CoefficientMat c;                                                                          \
            for (int a = 0; a < 3; ++a)                                                                            \
                for (int k = 0; k <= DEG; ++k) c(a, k) = coeffs[(i * 3 + a) * (DEG + 1) + k];                      \
            tr->emplace_back(T[i], c); /* trajectory.hpp:505, as learning_planner.hpp:216 does */                  \
        }                                                                                                          \
        return tr;                                                                                                 \
    }                                                                                                              \
    extern "C" void ref_traj##DEG##_destroy(void *p) { delete static_cast<Trajectory<DEG> *>(p); }                 \
    extern "C" double ref_traj##DEG##_cost(void *p, int order) { return static_cast<Trajectory<DEG> *>(p)->getTrajCost(order); } \
    extern "C" double ref_traj##DEG##_total_duration(void *p) { return static_cast<Trajectory<DEG> *>(p)->getTotalDuration(); } \
    extern "C" int ref_traj##DEG##_pieces(void *p) { return static_cast<Trajectory<DEG> *>(p)->getPieceNum(); }    \
    extern "C" void ref_traj##DEG##_eval(void *p, double t, double *pos, double *vel, double *acc, double *jer) {  \
        Trajectory<DEG> *tr = static_cast<Trajectory<DEG> *>(p);                                                   \
        const Eigen::Vector3d P = tr->getPos(t), V = tr->getVel(t), A = tr->getAcc(t), J = tr->getJer(t);          \
        for (int a = 0; a < 3; ++a) { pos[a] = P(a); vel[a] = V(a); acc[a] = A(a); jer[a] = J(a); }                \
    }                                                                                                              \
    extern "C" void ref_traj##DEG##_positions(void *p, double *out /* [N+1][3] */) {                               \
        Trajectory<DEG> *tr = static_cast<Trajectory<DEG> *>(p);                                                   \
        const Eigen::Matrix3Xd P = tr->getPositions();                                                             \
        for (int i = 0; i < P.cols(); ++i)                                                                         \
            for (int a = 0; a < 3; ++a) out[3 * i + a] = P(a, i);                                                  \
    }                                                                                                              \
    extern "C" int ref_traj##DEG##_locate(void *p, double *t_inout) { return static_cast<Trajectory<DEG> *>(p)->locatePieceIdx(*t_inout); } \
    /* max-rate members (trajectory.hpp:598-646; per piece :177-313) with the reference's own gcopter/root_finder.hpp */ \
    extern "C" double ref_traj##DEG##_max_vel_rate(void *p) { return static_cast<Trajectory<DEG> *>(p)->getMaxVelRate(); } \
    extern "C" double ref_traj##DEG##_max_acc_rate(void *p) { return static_cast<Trajectory<DEG> *>(p)->getMaxAccRate(); } \
    extern "C" int ref_traj##DEG##_check_max_vel_rate(void *p, double v) { return static_cast<Trajectory<DEG> *>(p)->checkMaxVelRate(v) ? 1 : 0; } \
    extern "C" int ref_traj##DEG##_check_max_acc_rate(void *p, double a) { return static_cast<Trajectory<DEG> *>(p)->checkMaxAccRate(a) ? 1 : 0; }

REF_TRAJ_API(5)
REF_TRAJ_API(7)
