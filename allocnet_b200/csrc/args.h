// args.h -- plain argument blocks shared by the host API (mincob.cu) and the kernels.
#pragma once
namespace mincob {

// reference codes, gcopter/lbfgs.hpp:135-184
enum {
    LBFGS_CONVERGENCE = 0,
    LBFGS_STOP = 1,
    LBFGS_CANCELED = 2,
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

struct DevParams {
    int kappa;
    double mu, w_pos, w_vel, w_acc, w_jerk, vmax2, amax2, jmax2, rho;
    double imu, ikap;   // 1/mu, 1/kappa (an fp64 division is ~30 instructions on the device)
    double amax2q;      // amax2 / 4 (exact): the sample loop tests half the acceleration (minco_device.cuh::sample_kinematics)
    int penalties;  // 0: energy-only fast path (all weights zero)
    int mapping;    // MINCOB_MAP_AUTO / _THROUGHPUT / _LATENCY (include/mincob.h)
    int freeze;     // 1: MINCOB_FLAG_FREEZE_TIMES -- durations stay as given (d/dtau = 0), waypoints only
    // L-BFGS (gcopter/lbfgs.hpp:15-129)
    int mem, past, max_iter, max_ls;
    double g_eps, delta, min_step, max_step, f_dec, s_curv, cautious, mach_prec;
};

// one batch of problems in the C-ABI layouts of include/mincob.h
struct BatchArgs {
    int B, N, K;
    const double *head, *tail, *hpolys;
    const int *hrows;
    // evaluate
    const double *x_in;
    double *f_out, *g_out;
    // optimize
    double *x, *coeffs, *T;
    int *status, *iters, *evals;
    int *counter;               // work queue head
    const int *ready;           // optional: problems [0, *ready) have arrived in device memory (chunked upload in flight)
    double *hist;               // optimize: (s, y) history scratch, one slab per resident group
    double *mult;               // optimize: block-solve multiplier scratch, one slab per resident block
    double *lpark;              // optimize: parked xp/gp/d, one slab per resident block
    int planes_in_smem;         // optimize: stage each problem's half-planes in shared memory
    unsigned long long *total_evals;  // optional: sum of evaluations (for the roofline numerator)
};

// MINCO building blocks (setParameters / getEnergy / ... / propogateGrad)
struct MincoArgs {
    int B, N;
    const double *head, *tail, *inPs, *ts;
    const double *gdC_in, *gdT_in;             // propagate
    double *coeffs_asc, *energy, *gdC, *gdT, *flat;  // forward
    double *gradByPoints, *gradByTimes;        // propagate
};

// sampled feasibility report of optimized trajectories (the job of Piece::getMaxVelRate / checkMaxAccRate,
// gcopter/trajectory.hpp:177-314, on a fixed grid instead of by root finding)
struct CheckArgs {
    int B, N, K, samples;          // `samples` sub-intervals per piece (samples + 1 points, both ends)
    const double *coeffs, *T;      // [B][N][3][2S] Trajectory order, [B][N]
    const double *hpolys;          // [B][N][K][4] or nullptr
    const int *hrows;
    double *out;                   // [B][4]: max |v|, max |a|, max |j|, max_k (n_k.p + d_k)  (<= 0: inside the corridor)
};

// exact maxima of |v|, |a|, |j| over optimized trajectories: the answers of Piece<D>::getMaxVelRate / getMaxAccRate
// and Trajectory<D>::getMaxVelRate / getMaxAccRate (gcopter/trajectory.hpp:177-273, 598-622), by bracketing and
// bisecting the stationary points of |p^(d)(t)|^2 instead of Sturm isolation
struct RateArgs {
    int B, N, grid;                // `grid` bracketing sub-intervals per piece
    const double *coeffs, *T;      // [B][N][3][2S] Trajectory order, [B][N]
    double *out;                   // [B][3]: max |v|, max |a|, max |j|
};

}  // namespace mincob
