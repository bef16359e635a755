/* ============================================================================
 * include/mincob.h -- C-ABI of the B200-native batched MINCO trajectory optimizer.
 *
 * Drop-in boundary for the trajectory back-end slot of AllocNet's planner.  The reference
 * has no FFI layer; the seam this library plugs into is the single in-process call
 *     qp_solver.solve(iniPVA, finPVA, hPolys, times, flatten_coffmats)
 *         src/planner/include/planner/learning_planner.hpp:196   (unflatten :201-233)
 * and the optimizer-side callback ABI
 *     lbfgs_evaluate_t / lbfgs_optimize   src/planner/include/gcopter/lbfgs.hpp:200-202, 434-440
 * The MINCO_S3NU / costFunctional API named by BASELINE.json:north_star is upstream GCOPTER
 * (minco.hpp, gcopter.hpp); it is NOT vendored in the reference and
 * is restated in SURVEY.md Appendix A/B.  Each entry point below cites what it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.  `*_device` variants take DEVICE
 *     pointers and enqueue on the handle's stream without synchronising; the plain variants
 *     take HOST pointers, copy in/out and return after the copies are complete (the caller may
 *     reuse or free its buffers at once).  Every entry point leaves the calling thread's current
 *     CUDA device as it found it.
 *   - all real data is fp64 (the reference path is double: lbfgs.hpp, trajectory.hpp).
 *   - return value: 0 on success, a negative MINCOB_E_* code otherwise; per-problem L-BFGS
 *     outcomes use the reference codes of gcopter/lbfgs.hpp:135-184 in the `status` array.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     MINCOB_E_CUDA (the oracle under oracle/ is test infrastructure, never linked here).
 *
 * Layouts (C order, B = problems, N = pieces, S = 3 (jerk, MINCO_S3NU) or 4 (snap, S4NU),
 *          n = N + 3(N-1) decision variables, K = half-plane rows stored per polytope)
 *   head, tail [B][S][3]     rows P,V,A[,J]; identical bytes to the Eigen col-major 3xS
 *                            iniPVA/finPVA of learning_planning.cpp:147-151
 *   hpolys     [B][N][K][4]  (nx,ny,nz,d) with n.p + d <= 0  (gcopter/geo_utils.hpp:41-42);
 *                            one polytope per piece as in planner/qp_solver.hpp:126,255-259.
 *                            Planner-form rows [n,b] (n.p <= b, learning_planner.hpp:293-299)
 *                            are converted by negating column 3.
 *   hrows      [B][N]        rows actually used per polytope (<= K); rest ignored
 *   x, g       [B][n]        x = [tau_0..tau_{N-1} ; q_1 .. q_{N-1} (xyz each)], T = forwardT(tau)
 *   coeffs     [B][N][3][2S] Trajectory<2S-1> order, k = 0 is the HIGHEST power
 *                            (gcopter/trajectory.hpp:79-83; flatten index idx = i*3*d + j*d + k of
 *                            planner/qp_solver.hpp:133 consumed at learning_planner.hpp:212)
 *   T          [B][N]        piece durations
 * ========================================================================== */
#ifndef MINCOB_H_
#define MINCOB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MINCOB_MAX_PIECES 32
#define MINCOB_MAX_ROWS 64     /* half-plane rows per polytope (the reference pads to 50: learning_planner.hpp:40,157-168) */
#define MINCOB_MAX_MEM 32
#define MINCOB_MAX_PAST 8

enum {
    MINCOB_OK = 0,
    MINCOB_E_INVALID = -1,   /* bad argument (null pointer, N out of range, S not 3/4, ...) */
    MINCOB_E_CUDA = -2,      /* CUDA runtime error or no device; see mincob_last_error() */
    MINCOB_E_STATE = -3,     /* call order (e.g. evaluate before set_problems) */
    MINCOB_E_NCCL = -4,      /* NCCL unavailable or failed */
    MINCOB_E_ALLOC = -5
};

/* Parameter block.  Penalty fields: upstream GCOPTER config (SURVEY.md Appendix B.1/B.2;
 * SmoothingEps in config/planner.yaml:15, MaxVelBox/MaxAccBox :17,:19, max_jerk in
 * network/utils/params.yaml:4).  L-BFGS fields: lbfgs_parameter_t, gcopter/lbfgs.hpp:15-129,
 * same names and meaning.  Field order is shared with oracle/oracle_capi.cpp:orc_params. */
typedef struct mincob_params {
    int32_t S;               /* 3: MINCO_S3NU (quintic, jerk energy); 4: MINCO_S4NU */
    int32_t kappa;           /* IntegralIntervs: trapezoid sub-intervals per piece */
    double mu;               /* SmoothingEps of smoothedL1, gcopter/firi.hpp:60-84 */
    double w_pos, w_vel, w_acc, w_jerk;
    double v_max, a_max, j_max;
    double rho;              /* WeightT */
    int32_t mem_size, past, max_iterations, max_linesearch;
    double g_epsilon, delta, min_step, max_step, f_dec_coeff, s_curv_coeff, cautious_factor, machine_prec;
    int32_t flags;           /* MINCOB_FLAG_* bits, 0 by default */
    int32_t mapping;         /* MINCOB_MAP_*: how mincob_optimize lays trajectories onto warps */
} mincob_params;

/* flags.
 * MINCOB_FLAG_FREEZE_TIMES: the durations handed in through x (tau = backwardT(T)) are DATA: only the inner
 *   waypoints are optimised (d/dtau = 0, T never moves).  This is the like-for-like replacement of the incumbent
 *   back-end call qp_solver.solve(iniPVA, finPVA, hPolys, times, ...) at planner/learning_planner.hpp:196, which
 *   keeps the network's time allocation fixed (planner/qp_solver.hpp:119-360).  Without it the optimizer also
 *   refines the durations (upstream GCOPTER behaviour).
 * MINCOB_FLAG_PLANNER_ROWS: half-plane rows are given as [n, b] meaning n.p <= b, the form LearningPlanner hands
 *   its back-end after the sign flip of planner/learning_planner.hpp:293-299, instead of GCOPTER's n.p + d <= 0
 *   (gcopter/geo_utils.hpp:41-42).  Read by mincob_set_problems / mincob_set_problems_device, which then keep a
 *   converted device copy of the rows. */
#define MINCOB_FLAG_FREEZE_TIMES 1
#define MINCOB_FLAG_PLANNER_ROWS 2
/* mapping (mincob_optimize only).
 * THROUGHPUT: one lane per piece, 32/LPT trajectories per warp (LPT = 5, 8, 16 or 32 lanes for N <= 5, 8, 16, 32).
 * LATENCY: "one warp per trajectory" -- the lane groups of a warp hold the same trajectory and split the tests of
 *   the penalty samples of every piece; fewer trajectories in flight, each evaluation ~1.3x shorter.  For single
 *   problems and small batches (needs kappa <= 31, otherwise THROUGHPUT runs).
 * AUTO: LATENCY when the batch has no more trajectories than the device holds resident warps, else THROUGHPUT.
 * The two mappings accumulate the active penalty samples in the same order with the same arithmetic: every output
 * is bit-identical in both, reproducible from run to run and independent of what else is in the batch. */
#define MINCOB_MAP_AUTO 0
#define MINCOB_MAP_THROUGHPUT 1
#define MINCOB_MAP_LATENCY 2

typedef struct mincob_ctx *mincob_handle;

/* ---- lifetime ------------------------------------------------------------------------- */
int mincob_default_params(mincob_params *out, int S);
int mincob_create(mincob_handle *out, const mincob_params *params, int device);
int mincob_destroy(mincob_handle h);
int mincob_set_params(mincob_handle h, const mincob_params *params);
/* Use an existing CUDA stream (cudaStream_t passed as void*); NULL = the handle's own (non-blocking) stream.
 * The legacy default stream is also NULL as a cudaStream_t: to run on it pass cudaStreamLegacy ((void*)0x1),
 * otherwise work on the handle's stream is not ordered with default-stream work of the caller. */
int mincob_set_stream(mincob_handle h, void *cuda_stream);
int mincob_synchronize(mincob_handle h);
const char *mincob_last_error(mincob_handle h);
const char *mincob_strerror(int code);
/* lbfgs_strerror of gcopter/lbfgs.hpp:724-800 for the per-problem status codes. */
const char *mincob_lbfgs_strerror(int status);
int mincob_version(void);

/* ---- problems: what LearningPlanner::callModel hands the back-end
 *      (iniPVA, finPVA, hPolys; learning_planner.hpp:140-196) for B problems ------------- */
int mincob_set_problems(mincob_handle h, int B, int N, int K, const double *head, const double *tail,
                        const double *hpolys, const int32_t *hrows);
int mincob_set_problems_device(mincob_handle h, int B, int N, int K, const double *head_d,
                               const double *tail_d, const double *hpolys_d, const int32_t *hrows_d);
/*      Overlapped form of mincob_set_problems for large batches: returns at once; the arrays are uploaded in problem
 *      order, in chunks, on a second stream, and the next mincob_optimize* call starts immediately -- its work queue
 *      waits per problem for a device-side arrival counter, so all but the first chunk of the upload is hidden behind
 *      the kernel.  The host arrays must be page-locked (mincob_host_alloc / mincob_host_register) and must not be
 *      modified until that optimize call has returned.  Any other entry point waits for the complete upload. */
int mincob_set_problems_async(mincob_handle h, int B, int N, int K, const double *head, const double *tail,
                              const double *hpolys, const int32_t *hrows);

/* ---- costFunctional (upstream gcopter.hpp) for every problem:
 *      the lbfgs_evaluate_t callback body (gcopter/lbfgs.hpp:200-202) as ONE kernel launch. */
int mincob_evaluate(mincob_handle h, const double *x, double *f, double *g);
int mincob_evaluate_device(mincob_handle h, const double *x_d, double *f_d, double *g_d);

/* ---- lbfgs::lbfgs_optimize (gcopter/lbfgs.hpp:434-717) on costFunctional for every problem,
 *      followed by getTrajectory (Appendix A.2).  x is in/out; any output pointer but x may be
 *      NULL.  status: LBFGS_CONVERGENCE(0) / LBFGS_STOP(1) / LBFGSERR_* per problem. */
int mincob_optimize(mincob_handle h, double *x, double *f, int32_t *status, int32_t *iters,
                    int32_t *evals, double *coeffs, double *T);
int mincob_optimize_device(mincob_handle h, double *x_d, double *f_d, int32_t *status_d, int32_t *iters_d,
                           int32_t *evals_d, double *coeffs_d, double *T_d);
/* Device time (ms, CUDA events on the handle's stream) and launch count of the last
 * evaluate/optimize call; synchronises. */
int mincob_last_kernel_ms(mincob_handle h, float *ms, int *launches);
/* MINCOB_MAP_THROUGHPUT / MINCOB_MAP_LATENCY: the mapping the last mincob_optimize* call launched. */
int mincob_last_mapping(mincob_handle h, int *mapping);

/* ---- MINCO_S3NU / MINCO_S4NU building blocks for B problems:
 *      setParameters + getCoeffs + getEnergy + getEnergyPartialGradBy{Coeffs,Times} ...
 *      inPs [B][N-1][3], ts [B][N]; coeffs_asc [B][2S*N][3] ascending powers (getCoeffs layout),
 *      gdC [B][2S*N][3], gdT [B][N], flat [B][N][3][2S] (getTrajectory).  NULL outputs skipped. */
int mincob_minco_forward(mincob_handle h, int B, int N, const double *head, const double *tail,
                         const double *inPs, const double *ts, double *coeffs_asc, double *energy,
                         double *gdC, double *gdT, double *flat);
/*      ... and propogateGrad (upstream spelling): partial dJ/dc, dJ/dT -> total dJ/dq, dJ/dT.
 *      gradByPoints [B][N-1][3], gradByTimes [B][N]. */
int mincob_minco_propagate(mincob_handle h, int B, int N, const double *head, const double *tail,
                           const double *inPs, const double *ts, const double *gdC, const double *gdT,
                           double *gradByPoints, double *gradByTimes);

/* Device-pointer forms of the two calls above (same layouts): enqueue one kernel on the handle's stream and
 * return; used by the differentiable layer allocnet_b200/autograd.py (forward = setParameters + getEnergy...,
 * backward = propogateGrad). */
int mincob_minco_forward_device(mincob_handle h, int B, int N, const double *head_d, const double *tail_d,
                                const double *inPs_d, const double *ts_d, double *coeffs_asc_d, double *energy_d,
                                double *gdC_d, double *gdT_d, double *flat_d);
int mincob_minco_propagate_device(mincob_handle h, int B, int N, const double *head_d, const double *tail_d,
                                  const double *inPs_d, const double *ts_d, const double *gdC_d, const double *gdT_d,
                                  double *gradByPoints_d, double *gradByTimes_d);

/* ---- feasibility report of optimized trajectories: what Piece::getMaxVelRate / getMaxAccRate /
 *      checkMaxVelRate (gcopter/trajectory.hpp:177-314) answer by root finding, on a fixed grid of
 *      `samples`+1 points per piece (both ends included), plus the largest corridor residual against
 *      the polytopes of the current mincob_set_problems call.
 *      coeffs [B][N][3][2S] and T [B][N] as returned by mincob_optimize;
 *      report [B][4] = max |v|, max |a|, max |j|, max_k (n_k.p + d_k)   (last entry <= 0: inside). */
int mincob_check_feasibility(mincob_handle h, const double *coeffs, const double *T, int samples, double *report);
int mincob_check_feasibility_device(mincob_handle h, const double *coeffs_d, const double *T_d, int samples,
                                    double *report_d);

/* ---- exact rate maxima of optimized trajectories: per trajectory, what Trajectory<D>::getMaxVelRate /
 *      getMaxAccRate (gcopter/trajectory.hpp:598-622; per piece :177-273, which isolate the roots of
 *      d/dt |p^(d)|^2 with RootFinder::solvePolynomial) return, plus the same for jerk.  The device brackets the
 *      stationary points on 128 sub-intervals per piece and bisects them.  checkMaxVelRate(v) / checkMaxAccRate(a)
 *      (:275-313, :624-646) are `rates[b][0] < v` / `rates[b][1] < a`.
 *      coeffs [B][N][3][2S] and T [B][N] as returned by mincob_optimize (B, N of the current
 *      mincob_set_problems call);  rates [B][3] = max |v|, max |a|, max |j|. */
int mincob_max_rates(mincob_handle h, const double *coeffs, const double *T, double *rates);
int mincob_max_rates_device(mincob_handle h, const double *coeffs_d, const double *T_d, double *rates_d);

/* ---- measured fp64 ceiling of the device: independent DFMA chains on every SM, timed with CUDA events on the
 *      handle's stream; TFLOP/s (2 flop per DFMA).  No counterpart in the reference (a CPU library); bench.py
 *      reports the optimize kernel's fp64 flop rate against it next to the HBM roofline. */
int mincob_measure_fp64_peak(mincob_handle h, double *tflops);

/* ---- multi-GPU: one process per GPU, problems block-partitioned, ONE all-gather of the solved
 *      coefficients (BASELINE.json north_star).  unique_id is the 128-byte ncclUniqueId made
 *      on rank 0 by mincob_nccl_unique_id and broadcast by the caller (torch.distributed). */
int mincob_nccl_unique_id(void *unique_id_128);
int mincob_comm_init(mincob_handle h, int nranks, int rank, const void *unique_id_128);
int mincob_allgather_device(mincob_handle h, const double *send_d, double *recv_d, int64_t count_per_rank);
int mincob_comm_destroy(mincob_handle h);
/*      HOST-pointer form of the sharded job (BASELINE.json config 5): optimize this rank's B problems
 *      (as mincob_optimize), all-gather the Trajectory-order coefficients of every rank, and copy
 *      them to coeffs_all [nranks*B][N][3][2S] (rank-major).  Without mincob_comm_init (one rank)
 *      it is mincob_optimize.  Every rank must call it with the same B. */
int mincob_optimize_sharded(mincob_handle h, double *x, double *f, int32_t *status, int32_t *iters,
                            int32_t *evals, double *coeffs_all, double *T);
/*      Same job and same collective, but only this rank's own coefficients [B][N][3][2S] are copied to the host
 *      (coeffs_local may be NULL); the gathered array of all ranks stays in device memory, where
 *      mincob_gathered_device returns it ([nranks*B][N][3][2S], rank-major; valid until the next sharded call).
 *      A job in which one rank consumes everything calls mincob_optimize_sharded there and this on the others. */
int mincob_optimize_sharded_local(mincob_handle h, double *x, double *f, int32_t *status, int32_t *iters,
                                  int32_t *evals, double *coeffs_local, double *T);
int mincob_gathered_device(mincob_handle h, const double **coeffs_all_d, int64_t *count);

/* ---- pinned host memory for the host-pointer entry points (cudaHostAlloc / cudaFreeHost), so a
 *      C++ caller such as LearningPlanner needs no CUDA headers of its own. */
int mincob_host_alloc(void **out, uint64_t bytes);
int mincob_host_free(void *p);
/*      ... and page-locking of memory the caller owns (cudaHostRegister / cudaHostUnregister), e.g. a shared-memory
 *      segment mapped by all ranks of a node: every rank's results then land directly in the consumer's memory. */
int mincob_host_register(void *p, uint64_t bytes);
int mincob_host_unregister(void *p);

#ifdef __cplusplus
}
#endif
#endif /* MINCOB_H_ */
