// ============================================================================
// lbfgs_device.cuh -- persistent batched L-BFGS driver, one LPT-lane group per trajectory.
//
// Control flow restates lbfgs::lbfgs_optimize (src/planner/include/gcopter/lbfgs.hpp:434-717)
// and line_search_lewisoverton (:276-384) of the reference: same tests, same order, same
// return codes (:135-184).  It is written as a per-group state machine with exactly ONE
// cost-functional call site per loop trip so that the 32/LPT groups of a warp, each somewhere
// else in its own line search, execute the expensive evaluation convergently.  A group that
// finishes writes its results (x, f, status, iterations, evaluations, Trajectory-order
// coefficients, durations) and pulls the next problem from a global counter, so there are no
// half-empty launches and no per-iteration HBM traffic for optimizer state.
// Where the state lives: x, g, xp, gp, d in registers (lane i owns tau_i and q_i; xp, gp, d and
// the group's scalars are parked in an L2-resident slab / shared memory while the cost functional
// runs, which is what lets three blocks share an SM); the problem's half-planes, head/tail and
// the small rings (1/y.s, past f) in shared memory; the (s, y) history in a global slab per
// resident group that never leaves L2.  DESIGN.md section 3 has the table.
// ============================================================================
#pragma once
#include "minco_device.cuh"

#ifndef MINCOB_PARK
#define MINCOB_PARK 1
#endif
#ifndef MINCOB_LOCKSTEP
#define MINCOB_LOCKSTEP 0
#endif
#ifndef MINCOB_HDEP
#define MINCOB_HDEP 2   // (s, y) pairs of the two-loop recursion in flight ahead of their use (MEM > 0 kernels)
#endif
// MINCOB_STRICT: the decision arithmetic of gcopter/lbfgs.hpp written exactly as the reference writes it (quotients,
// square roots, divisions by y.s) instead of the division-free forms of the default build, which are equal in exact
// arithmetic but can round a decision differently within an ulp of a threshold.  build.py also builds this variant
// (allocnet_b200/libmincob_strict.so); tests/test_gpu_parity.py::test_division_free_decisions_match_reference_forms
// measures how often the two builds decide differently.
#ifndef MINCOB_STRICT
#define MINCOB_STRICT 0
#endif
#ifndef MINCOB_MINB
#define MINCOB_MINB 2   // resident blocks per SM the optimize kernel is compiled for (register cap)
#endif

namespace mincob {



// problem p as it lies in global memory (C-ABI layout of include/mincob.h)
template <int S>
__device__ __forceinline__ ProblemView view_global(const BatchArgs &a, int p, int lig) {
    ProblemView pv;
    pv.head = a.head + (size_t)p * S * 3;
    pv.tail = a.tail + (size_t)p * S * 3;
    const bool have = a.hpolys && a.hrows && a.K > 0;
    const int l = lig < a.N ? lig : 0;
    pv.planes = have ? a.hpolys + ((size_t)p * a.N + l) * a.K * 4 : nullptr;
    pv.rstride = 4;
    pv.rows = have ? min(a.hrows[(size_t)p * a.N + l], a.K) : 0;
    return pv;
}

// lane `lig` owns v[0] = tau_lig (lanes < N) and v[1..3] = q_lig (lanes 1..N-1); everything else 0.
template <int LPT>
__device__ __forceinline__ void load_x(const double *x, int N, int lig, double (&v)[4]) {
    v[0] = (lig < N) ? x[lig] : 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) v[1 + a] = (lig >= 1 && lig < N) ? x[N + 3 * (lig - 1) + a] : 0.0;
}
template <int LPT>
__device__ __forceinline__ void store_x(double *x, int N, int lig, const double (&v)[4]) {
    if (lig < N) x[lig] = v[0];
    if (lig >= 1 && lig < N) {
#pragma unroll
        for (int a = 0; a < 3; ++a) x[N + 3 * (lig - 1) + a] = v[1 + a];
    }
}
template <int LPT>
__device__ __forceinline__ double gdot(unsigned m, const double (&a)[4], const double (&b)[4]) {
    return group_sum<LPT>(m, a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]);
}
template <int LPT>
__device__ __forceinline__ double ginf(unsigned m, const double (&a)[4]) {
    return group_max<LPT>(m, fmax(fmax(fabs(a[0]), fabs(a[1])), fmax(fabs(a[2]), fabs(a[3]))));
}

// setParameters + getTrajectory at x: writes Trajectory-order coefficients and durations.
template <int S, int LPT>
__device__ __noinline__ void emit_trajectory(unsigned, int lig, int N, int rounds, const ProblemView &pv,
                                             const double (&xv)[4], double *coeffs, double *Tout) {
    constexpr int D = 2 * S, b = S - 1;
    constexpr unsigned mask = 0xffffffffu;  // called by whole warps only
    const bool active = lig < N;
    double P0[3], P1[3], hd[b][3], td[b][3];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        const double qn = sh_dn<LPT>(mask, xv[1 + x], 1);
        P0[x] = (lig == 0) ? pv.head[x] : xv[1 + x];
        P1[x] = (lig == N - 1) ? pv.tail[x] : qn;
    }
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            hd[a][x] = (lig == 0) ? pv.head[(a + 1) * 3 + x] : 0.0;
            td[a][x] = (lig == N - 1) ? pv.tail[(a + 1) * 3 + x] : 0.0;
        }
    const double T = active ? forward_t(xv[0]) : 1.0;
    SplineReg<S, LPT> sp;
    double chat[D][3];
    spline_solve<S, LPT>(mask, lig, N, rounds, T, P0, P1, hd, td, sp, chat);
    if (active) {
        if (coeffs) {
            // Trajectory<2S-1>: [piece][axis][k], k = 0 highest power (gcopter/trajectory.hpp:79-83)
#pragma unroll
            for (int x = 0; x < 3; ++x)
#pragma unroll
                for (int k = 0; k < D; ++k) coeffs[((size_t)lig * 3 + x) * D + k] = sp.c[D - 1 - k][x];
        }
        if (Tout) Tout[lig] = T;
    }
}

enum { PH_FETCH = 0, PH_FIRST = 1, PH_LS = 2, PH_IDLE = 3 };

// doubles of shared memory one group needs (host and device must agree): the plane stage, then the small part
// (head/tail, alpha / y.s / past-f rings, parking).  The dummy group of a warp (LPT not a power of two) only
// gets a small part.
__host__ __device__ inline int optimize_small_doubles(int S, int m, int past, int lpt) {
    int small = 2 * S * 3 + 2 * m + (past > 0 ? past : 1) + 12 + 4 * lpt;   // + parking: scalars (7 doubles, 8 ints), d
    return (small + 3) & ~3;                       // keep every group's plane stage 32-byte aligned
}
__host__ __device__ inline int optimize_group_doubles(int S, int N, int K, int m, int past, int planes_in_smem, int lpt) {
    return (planes_in_smem ? N * K * 4 : 0) + optimize_small_doubles(S, m, past, lpt);
}

// Control flow of one trip (all 32 lanes of the warp walk it together; the 32/LPT groups differ only
// in predicates, so every shuffle uses the full mask and no branch encloses a shuffle):
//   fetch (+ stage the problem in shared memory) -> costFunctional -> reductions -> per-group scalar
//   decisions (lbfgs.hpp tests, in the reference's order) -> history update + two-loop recursion ->
//   line-search entry -> retire.
// Memory: x, g, xp, gp, d in registers (lane i owns tau_i and q_i).  Shared memory per group: the
// problem's half-planes, transposed to plane-major so that one row index is one conflict-free
// 32*N-byte read for the group (they are read 3x17x16 times per evaluation, ~400 evaluations per
// problem), head/tail states, alpha / y.s / past-f rings.  The (s, y) history (touched once per
// iteration) lives in a global scratch slab per resident group, which stays in L2.
// MEM > 0: the history depth is the compile-time constant MEM (== P.mem, the launcher checks): the two-loop
// recursion is unrolled over registers, the first MINCOB_HDEP history pairs are requested right after the
// evaluation (their L2 latency is covered by the reductions and the scalar decisions) and the others
// MINCOB_HDEP steps ahead of their use.  MEM == 0: any depth, rolled loops with two slots in flight.
// PSM: 0 half-planes read from global memory, 1 staged in shared memory.
// REP: latency mapping, "one warp per trajectory" (BASELINE.json north_star): the GPW groups of a warp fetch the SAME
// problem and run the same state machine on bitwise identical state; only the penalty samples are split among them
// (cost_functional<.., REP>).  Used when the batch is too small to fill the device with one group per problem.
// FRZ: fixed-time mode (MINCOB_FLAG_FREEZE_TIMES, the call the reference makes: learning_planner.hpp:196) as its own
// instantiation: the block factorisation of a problem is computed by its first evaluation and reused by all later ones,
// and no time gradient is formed.  Bit-identical to the generic kernel run with P.freeze
// (tests/test_gpu_parity.py::test_fixed_time_kernel_equals_generic_kernel).
template <int S, int LPT, int THREADS, int PSM, int MEM, bool REP, bool FRZ = false>
__global__ void __launch_bounds__(THREADS, MINCOB_MINB) optimize_kernel(const DevParams P, const BatchArgs a) {
    constexpr bool PSMEM = PSM != 0;
    constexpr unsigned FULL = 0xffffffffu;
    using LN = Lanes<LPT>;
    constexpr int SPB = (THREADS / 32) * LN::SPW;   // group slots per block (real groups + dummy groups)
    const int lig = LN::lig();
    const int gib = LN::slot();
    const int N = a.N, K = a.K, n = N + 3 * (N - 1), m = MEM > 0 ? MEM : P.mem, past = P.past;
    const int rounds = N > 2 ? N - 2 : 0;   // lane-to-lane sweeps of the block solve (warp-uniform)

    extern __shared__ __align__(32) double smem[];
    // real groups first (plane stage + small part each), then one small part per dummy group
    constexpr int GPBK = (THREADS / 32) * LN::GPW;
    const int gdoubles = optimize_group_doubles(S, N, K, m, past, PSMEM ? 1 : 0, LPT);
    double *planes_s = smem + (size_t)(LN::real() ? LN::gib() : 0) * gdoubles;   // [K][N][4] when PSMEM
    double *ht = LN::real() ? planes_s + (PSMEM ? N * K * 4 : 0)                  // head [S][3], tail [S][3]
                            : smem + (size_t)GPBK * gdoubles + (size_t)(threadIdx.x >> 5) * optimize_small_doubles(S, m, past, LPT);
    double *alpha = ht + 2 * S * 3;
    double *ysv = alpha + m;
    double *pf = ysv + m;
    double *park_d = pf + (past > 0 ? past : 1);             // L-BFGS scalars of the group while costFunctional runs
    int *park_i = reinterpret_cast<int *>(park_d + 8);
    double *park_dir = park_d + 12 + lig;                     // d of this lane (needed first after the evaluation): [4][LPT]
    // history slab of this group: [m][LPT][8] = (s0..s3, y0..y3) of lane lig in slot j
    double *hist = a.hist + ((size_t)blockIdx.x * SPB + gib) * ((size_t)m * LPT * 8) + (size_t)lig * 8;

    // block-solve multipliers of the current evaluation: slot i of thread t at mult[(block*NM + i)*THREADS + t]
    GlobalStore mstore;
    mstore.p = a.mult + (size_t)blockIdx.x * SplineReg<S, LPT>::NM * THREADS + threadIdx.x;
    mstore.stride = THREADS;
    GlobalStore lstore;   // xp, gp of this lane while the cost functional runs (8 of 12 slots used)
    lstore.p = a.lpark + (size_t)blockIdx.x * 12 * THREADS + threadIdx.x;
    lstore.stride = THREADS;

#ifdef MINCOB_TIMING
    if (threadIdx.x == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMin(a.total_evals + 3, t);
    }
#endif
    int phase = LN::real() ? PH_FETCH : PH_IDLE, prob = 0;   // the dummy group (LPT not a power of two) never owns a problem
    if (!LN::real()) {
        for (int i = lig; i < 2 * S * 3; i += LPT) ht[i] = 0.0;
    }
    ProblemView pv;
    pv.head = ht; pv.tail = ht + S * 3; pv.planes = nullptr; pv.rstride = 4; pv.rows = 0;
    double x[4] = {0, 0, 0, 0}, g[4], xp[4] = {0, 0, 0, 0}, gp[4] = {0, 0, 0, 0}, d[4] = {0, 0, 0, 0};
    double fx = 0.0, stp = 0.0, finit = 0.0, dgtest = 0.0, dstest = 0.0, lo = 0.0, hi = 0.0;
    int count = 0, k = 0, end = 0, bound = 0, evals = 0;
    bool bracketed = false, touched = false;
    unsigned long long my_evals = 0ull;

    for (;;) {
        // ---- fetch ---------------------------------------------------------------------------
        if (__any_sync(FULL, phase == PH_FETCH)) {
            const bool want = phase == PH_FETCH;
            int p = 0;
            if (REP) {   // the replicas of a warp always finish together: one ticket for the warp
                if (want && LN::lane() == 0) p = atomicAdd(a.counter, 1);
                p = __shfl_sync(FULL, p, 0);
            } else {
                if (want && lig == 0) p = atomicAdd(a.counter, 1);
                p = group_first<LPT>(FULL, p);
            }
            if (a.ready != nullptr) {
                // the batch is still being uploaded in problem order (mincob_set_problems_async): wait for this problem
                if (want && p < a.B) {
                    while (*(volatile const int *)a.ready <= p) __nanosleep(200);
                }
                __syncwarp();
            }
            if (want) {
                if (p >= a.B) {
#ifdef MINCOB_TIMING
                    if (lig == 0) {   // experiment: when did the work queue run dry / how many groups idle over time
                        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                        atomicMin(a.total_evals + 1, t);
                        const unsigned long long t0 = *(volatile unsigned long long *)(a.total_evals + 1);
                        const unsigned long long bin = (t - t0) / 1000000ull;   // 1 ms bins after the queue ran dry
                        if (bin < 64) atomicAdd(a.total_evals + 4 + bin, 1ull);
                    }
#endif
                    phase = PH_IDLE; prob = 0;
                    for (int i = lig; i < 2 * S * 3; i += LPT) ht[i] = 0.0;
                    pv.planes = nullptr; pv.rows = 0;
                } else {
                    prob = p;
                    load_x<LPT>(a.x + (size_t)p * n, N, lig, x);
                    phase = PH_FIRST;
                    for (int i = lig; i < 2 * S * 3; i += LPT)
                        ht[i] = i < S * 3 ? a.head[(size_t)p * S * 3 + i] : a.tail[(size_t)p * S * 3 + i - S * 3];
                    const bool have = a.hpolys && a.hrows && K > 0;
                    pv.rows = (have && lig < N) ? min(a.hrows[(size_t)p * N + lig], K) : 0;
                    if (PSMEM) {
                        // [N][K][4] (global, one contiguous 32*N*K-byte block) -> [K][N][4] (shared)
                        const double2 *src = reinterpret_cast<const double2 *>(a.hpolys + (size_t)p * N * K * 4);
                        double2 *dst = reinterpret_cast<double2 *>(planes_s);
                        for (int r = lig; r < N * K; r += LPT) {
                            const int piece = r / K, kk = r - piece * K;
                            const double2 lo2 = __ldg(src + 2 * r), hi2 = __ldg(src + 2 * r + 1);
                            dst[2 * (kk * N + piece)] = lo2;
                            dst[2 * (kk * N + piece) + 1] = hi2;
                        }
                        pv.planes = have ? planes_s + (lig < N ? lig : 0) * 4 : nullptr;
                        pv.rstride = 4 * N;
                    } else {
                        pv.planes = have ? a.hpolys + ((size_t)p * N + (lig < N ? lig : 0)) * K * 4 : nullptr;
                        pv.rstride = 4;
                    }
                }
            }
            __syncwarp();
        }
#if MINCOB_LOCKSTEP
        // the warps of a block walk the (cache-exceeding) evaluation code together, sharing its fetches
        if (__syncthreads_and(phase == PH_IDLE)) break;
#else
        if (__all_sync(FULL, phase == PH_IDLE)) break;
#endif

        // ---- costFunctional --------------------------------------------------------------------
        double xq[3] = {x[1], x[2], x[3]}, gq[3];
#if MINCOB_PARK
        // Nothing of the optimizer state is needed while the cost functional runs: park xp, gp, d in the
        // lane-strided global slab and the group's scalars in shared memory, so that they do not hold
        // ~60 registers across the penalty loop (which is what decides how many warps fit on an SM).
        {
#pragma unroll
            for (int i = 0; i < 4; ++i) { lstore.put(i, xp[i]); lstore.put(4 + i, gp[i]); park_dir[i * LPT] = d[i]; }
            if (lig == 0) {
                park_d[0] = fx; park_d[1] = stp; park_d[2] = finit; park_d[3] = dgtest; park_d[4] = dstest;
                park_d[5] = lo; park_d[6] = hi;
                park_i[0] = count; park_i[1] = k; park_i[2] = end; park_i[3] = bound; park_i[4] = evals;
                park_i[5] = (bracketed ? 1 : 0) | (touched ? 2 : 0); park_i[6] = prob;
            }
            __syncwarp();
        }
#endif
#if MINCOB_PARK
        auto unpark = [&]() {
#pragma unroll
            for (int i = 0; i < 4; ++i) { xp[i] = lstore.get(i); gp[i] = lstore.get(4 + i); d[i] = park_dir[i * LPT]; }
        };
        // fixed-time kernel: a newly fetched problem anywhere in the warp makes the whole warp factorise (the other
        // groups recompute and store the multipliers they already hold)
        const bool refac = !FRZ || __any_sync(FULL, phase == PH_FIRST);
        const double f = cost_functional<S, LPT, PSM, GlobalStore, REP, decltype(unpark), FRZ>(P, FULL, lig, phase == PH_IDLE ? 0 : N, rounds, pv, mstore, x[0], xq, g[0], gq, unpark, refac);
#else
        const bool refac = !FRZ || __any_sync(FULL, phase == PH_FIRST);
        const double f = cost_functional<S, LPT, PSM, GlobalStore, REP, NoHook, FRZ>(P, FULL, lig, phase == PH_IDLE ? 0 : N, rounds, pv, mstore, x[0], xq, g[0], gq, NoHook(), refac);
#endif
        g[1] = gq[0]; g[2] = gq[1]; g[3] = gq[2];
#if MINCOB_PARK
        {
            fx = park_d[0]; stp = park_d[1]; finit = park_d[2]; dgtest = park_d[3]; dstest = park_d[4];
            lo = park_d[5]; hi = park_d[6];
            count = park_i[0]; k = park_i[1]; end = park_i[2]; bound = park_i[3]; evals = park_i[4];
            bracketed = park_i[5] & 1; touched = (park_i[5] & 2) != 0; prob = park_i[6];
        }
#endif
        // (s, y) pairs the two-loop recursion will want if this trial point is accepted: after the update the
        // i-th newest pair (i >= 1) sits in slot end - i of the ring as it is now; pair 0 is the one about to be
        // computed.  The first HDEP of them are requested here, so that the reductions and scalar decisions
        // below cover their L2 latency; the rest follow HDEP steps ahead of their use (all of them at once
        // would not fit in registers next to x, g, xp, gp, d).  Slots never written yet (bound < m) hold
        // garbage that the `i < bound` predicates keep out.
        struct Pair { double4 s, y; double rys; };
        constexpr int HDEP = MEM > 1 ? (MINCOB_HDEP < MEM - 1 ? MINCOB_HDEP : MEM - 1) : 0;
        const int end0 = end;
        auto fetch_nth = [&](int i) {          // i-th newest pair after the update, 1 <= i < MEM
            int jj = end0 - i;
            jj = jj < 0 ? jj + MEM : jj;
            const double4 *slot = reinterpret_cast<const double4 *>(hist + (size_t)jj * LPT * 8);
            Pair p;
            p.s = slot[0]; p.y = slot[1]; p.rys = ysv[jj];
            return p;
        };
        Pair hq[MEM > 1 ? MEM : 2];            // hq[i] = i-th newest (hq[0] = the new pair)
        if (MEM > 1) {
#pragma unroll
            for (int i = 1; i <= HDEP; ++i) hq[i] = fetch_nth(i);
        }
#ifdef MINCOB_HIST_PREFETCH
        if (MEM == 0) {   // experiment: rolled two-loop recursion, history lines pulled towards L1 ahead of it
#pragma unroll 1
            for (int j = 0; j < m; ++j) {
                const double *sl = hist + (size_t)j * LPT * 8;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(sl));
            }
        }
#endif

        // ---- reductions every group may need ------------------------------------------------------
        // |g|_inf and |x|_inf only feed the g_epsilon test (lbfgs.hpp:531, :600), which can never fire for g_epsilon = 0
        // (the default here and upstream): a warp-uniform branch then skips both reductions
        double gn = 1.0, xn = 1.0;
        if (MINCOB_STRICT || P.g_eps > 0.0) { gn = ginf<LPT>(FULL, g); xn = ginf<LPT>(FULL, x); }
        const double gg = gdot<LPT>(FULL, g, g);
        const double gd = gdot<LPT>(FULL, g, d);
        const int kslot = (past > 0) ? k % past : 0;   // slot of the `past` ring this iteration reads, then overwrites
        const double pfk = (past > 0) ? pf[kslot] : 0.0;
        __syncwarp();

        // ---- per-group scalar decisions (no shuffles below until the next section) -------------
        int finish = 0, ret = 0;
        bool start_ls = false, upd = false, write_pf = false;
        int pf_slot = 0;
        if (phase == PH_FIRST) {            // lbfgs.hpp:518-545
            evals = 1;
            fx = f;
            write_pf = true;                // pf(0) = fx
            k = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = -g[i];
#if MINCOB_STRICT
            if (gn / fmax(1.0, xn) < P.g_eps) {   // lbfgs.hpp:531 as written
#else
            if (gn < P.g_eps * fmax(1.0, xn)) {   // gnorm / max(1, xnorm) < g_epsilon (lbfgs.hpp:531), without the fp64 division
#endif
                ret = LBFGS_CONVERGENCE; finish = 1;
            } else {
#if MINCOB_STRICT
                stp = 1.0 / sqrt(gg);           // step = 1.0 / d.norm() (lbfgs.hpp:543)
#else
                stp = rsqrt(gg);                // 1 / |g| (lbfgs.hpp:543), one rounding instead of two
#endif
                k = 1; end = 0; bound = 0;
                start_ls = true;
            }
        } else if (phase == PH_LS) {        // one trial point of line_search_lewisoverton (:311-384)
            ++count; ++evals;
            fx = f;
            int fail = 0;
            bool done = false;
            if (isinf(f) || isnan(f)) {
                fail = LBFGSERR_INVALID_FUNCVAL;
            } else if (f > finit + stp * dgtest) {
                hi = stp; bracketed = true;
            } else if (gd < dstest) {
                lo = stp;
            } else {
                done = true;
            }
            if (!fail && !done) {
                if (P.max_ls <= count) fail = LBFGSERR_MAXIMUMLINESEARCH;
                else if (bracketed && (hi - lo) < P.mach_prec * hi) fail = LBFGSERR_WIDTHTOOSMALL;
                else {
                    stp = bracketed ? 0.5 * (lo + hi) : stp * 2.0;
                    if (stp < P.min_step) fail = LBFGSERR_MINIMUMSTEP;
                    else if (stp > P.max_step) {
                        if (touched) fail = LBFGSERR_MAXIMUMSTEP;
                        else { touched = true; stp = P.max_step; }
                    }
                }
                if (!fail) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = xp[i] + stp * d[i];
                }
            }
            if (fail) {                     // lbfgs.hpp:570-577: revert to the last good iterate
#pragma unroll
                for (int i = 0; i < 4; ++i) { x[i] = xp[i]; g[i] = gp[i]; }
                ret = fail; finish = 1;
            } else if (done) {              // lbfgs.hpp:580-640
#if MINCOB_STRICT
                if (gn / fmax(1.0, xn) < P.g_eps) { ret = LBFGS_CONVERGENCE; finish = 1; }
#else
                if (gn < P.g_eps * fmax(1.0, xn)) { ret = LBFGS_CONVERGENCE; finish = 1; }
#endif
                if (!finish && past > 0) {
                    // |pf - fx| / max(1, |fx|) < delta (lbfgs.hpp:610-614) as a product: an fp64 division is ~35 instructions
#if MINCOB_STRICT
                    if (past <= k && fabs(pfk - fx) / fmax(1.0, fabs(fx)) < P.delta) { ret = LBFGS_STOP; finish = 1; }
#else
                    if (past <= k && fabs(pfk - fx) < P.delta * fmax(1.0, fabs(fx))) { ret = LBFGS_STOP; finish = 1; }
#endif
                    if (!finish) { write_pf = true; pf_slot = kslot; }
                }
                if (!finish && P.max_iter != 0 && P.max_iter <= k) { ret = LBFGSERR_MAXIMUMITERATION; finish = 1; }
                if (!finish) upd = true;
            }
        }
        if (write_pf && lig == 0) pf[pf_slot] = fx;

        // ---- new iterate: (s, y) pair, cautious update, two-loop recursion (lbfgs.hpp:642-709) ----
        if (__any_sync(FULL, upd)) {
            double sv[4], yv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { sv[i] = x[i] - xp[i]; yv[i] = g[i] - gp[i]; }
            const double ys = gdot<LPT>(FULL, yv, sv), yy = gdot<LPT>(FULL, yv, yv);
            const double ss = gdot<LPT>(FULL, sv, sv), gpgp = gdot<LPT>(FULL, gp, gp);
#if MINCOB_STRICT
            const double iys = ys;         // keep y.s itself and divide where lbfgs.hpp:676-701 divides
#else
            const double iys = 1.0 / ys;   // the two-loop recursion only ever divides by y.s
#endif
            bool two = false;
            if (upd) {
                ++k;
                double4 *slot = reinterpret_cast<double4 *>(hist + (size_t)end * LPT * 8);
                slot[0] = make_double4(sv[0], sv[1], sv[2], sv[3]);
                slot[1] = make_double4(yv[0], yv[1], yv[2], yv[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) d[i] = -g[i];
                if (lig == 0) ysv[end] = iys;
                // ys > cautious * ss * |gp| (lbfgs.hpp:655), squared so that no fp64 square root is needed (the right side is >= 0)
#if MINCOB_STRICT
                two = ys > P.cautious * ss * sqrt(gpgp);       // lbfgs.hpp:655 as written
#else
                two = ys > 0.0 && ys * ys > (ss * P.cautious) * (ss * P.cautious) * gpgp;
#endif
                if (two) {
                    ++bound;
                    bound = m < bound ? m : bound;
                    end = (end + 1 == m) ? 0 : end + 1;
                }
                stp = 1.0;
                start_ls = true;
            }
            __syncwarp();
            if (MEM > 0) {
                // HKEEP: pairs >= HKEEP stay in registers from the backward pass for the forward pass (which
                // starts with the oldest); the younger ones are requested again HDEP steps before their turn.
                constexpr int HKEEP = MEM - HDEP > 1 ? MEM - HDEP : 1;
                hq[0].s = make_double4(sv[0], sv[1], sv[2], sv[3]);
                hq[0].y = make_double4(yv[0], yv[1], yv[2], yv[3]);
                hq[0].rys = iys;
                double al[MEM > 0 ? MEM : 1];
                // backward pass, newest pair first (lbfgs.hpp:676-687)
#pragma unroll
                for (int i = 0; i < MEM; ++i) {
                    if (i >= 1 && i + HDEP < MEM) hq[i + HDEP] = fetch_nth(i + HDEP);
                    const bool on = two && i < bound;
                    const double sj[4] = {hq[i].s.x, hq[i].s.y, hq[i].s.z, hq[i].s.w};
                    const double aj = MINCOB_STRICT ? gdot<LPT>(FULL, sj, d) / hq[i].rys : gdot<LPT>(FULL, sj, d) * hq[i].rys;
                    al[i] = aj;
                    if (on) { d[0] -= aj * hq[i].y.x; d[1] -= aj * hq[i].y.y; d[2] -= aj * hq[i].y.z; d[3] -= aj * hq[i].y.w; }
                }
                if (two) {
                    const double sc0 = ys / yy;
#pragma unroll
                    for (int u = 0; u < 4; ++u) d[u] *= sc0;
                }
                // forward pass, oldest pair first (:691-701)
                Pair hr[MEM > 1 ? MEM : 2];        // re-requested young pairs, hr[i] = i-th newest (1 <= i < HKEEP)
#pragma unroll
                for (int i = MEM - 1; i >= 0; --i) {
                    if (i - HDEP >= 1) hr[i - HDEP] = fetch_nth(i - HDEP);
                    const Pair &h = (i >= HKEEP || i == 0) ? hq[i] : hr[i];
                    const bool on = two && i < bound;
                    const double yj[4] = {h.y.x, h.y.y, h.y.z, h.y.w};
                    const double beta = MINCOB_STRICT ? gdot<LPT>(FULL, yj, d) / h.rys : gdot<LPT>(FULL, yj, d) * h.rys;
                    if (on) {
                        const double cf = al[i] - beta;
                        d[0] += cf * h.s.x; d[1] += cf * h.s.y; d[2] += cf * h.s.z; d[3] += cf * h.s.w;
                    }
                }
            } else {
                const int nb = __reduce_max_sync(FULL, two ? bound : 0);
                // The history sits in L2 (a 4 KB slab per trajectory, too big for what shared memory is left),
                // every address of the recursion is known before it starts, and each step is a short dependent
                // chain: keep two slots in flight ahead of the one being consumed.
                // Backward pass, step i uses slot end-1-i (newest first); step 0 is the pair just computed.
                auto slot_at = [&](int i) {          // i-th newest slot of this group (i < m)
                    int jj = end - 1 - i;
                    return jj < 0 ? jj + m : jj;
                };
                auto fetch = [&](int jj) {
                    const double4 *slot = reinterpret_cast<const double4 *>(hist + (size_t)jj * LPT * 8);
                    Pair p;
                    p.s = slot[0]; p.y = slot[1];
                    return p;
                };
                Pair cur, n1;
                cur.s = make_double4(sv[0], sv[1], sv[2], sv[3]);
                cur.y = make_double4(yv[0], yv[1], yv[2], yv[3]);
                n1 = fetch(slot_at(nb > 1 ? 1 : 0));
#pragma unroll 1
                for (int i = 0; i < nb; ++i) {
                    const bool on = two && i < bound;
                    const Pair n2 = fetch(slot_at(i + 2 < nb ? i + 2 : nb - 1));
                    const int j = slot_at(i);
                    const double sj[4] = {cur.s.x, cur.s.y, cur.s.z, cur.s.w};
                    const double aj = MINCOB_STRICT ? gdot<LPT>(FULL, sj, d) / ysv[j] : gdot<LPT>(FULL, sj, d) * ysv[j];
                    if (on) {
                        if (lig == 0) alpha[j] = aj;
                        d[0] -= aj * cur.y.x; d[1] -= aj * cur.y.y; d[2] -= aj * cur.y.z; d[3] -= aj * cur.y.w;
                    }
                    cur = n1; n1 = n2;
                }
                __syncwarp();
                if (two) {
                    const double sc0 = ys / yy;
#pragma unroll
                    for (int u = 0; u < 4; ++u) d[u] *= sc0;
                }
                // Forward pass, step i uses this group's (bound-1-i)-th newest slot (oldest first).
                auto fwd_at = [&](int i) { return slot_at(bound - 1 - i > 0 ? bound - 1 - i : 0); };
                cur = fetch(fwd_at(0));
                n1 = fetch(fwd_at(1));
#pragma unroll 1
                for (int i = 0; i < nb; ++i) {
                    const bool on = two && i < bound;
                    const Pair n2 = fetch(fwd_at(i + 2));
                    const int j = fwd_at(i);
                    const double yj[4] = {cur.y.x, cur.y.y, cur.y.z, cur.y.w};
                    const double beta = MINCOB_STRICT ? gdot<LPT>(FULL, yj, d) / ysv[j] : gdot<LPT>(FULL, yj, d) * ysv[j];
                    if (on) {
                        const double cf = alpha[j] - beta;
                        d[0] += cf * cur.s.x; d[1] += cf * cur.s.y; d[2] += cf * cur.s.z; d[3] += cf * cur.s.w;
                    }
                    cur = n1; n1 = n2;
                }
            }
        }

        // ---- entry of line_search_lewisoverton (lbfgs.hpp:276-310) -------------------------------
        if (__any_sync(FULL, start_ls)) {
            if (start_ls) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { xp[i] = x[i]; gp[i] = g[i]; }
            }
            const double dginit = gdot<LPT>(FULL, gp, d);
            if (start_ls) {
                int fail = 0;
                if (!(stp > 0.0)) fail = LBFGSERR_INVALIDPARAMETERS;
                else if (0.0 < dginit) fail = LBFGSERR_INCREASEGRADIENT;
                if (fail) {
                    ret = fail; finish = 1;
                } else {
                    finit = fx;
                    dgtest = P.f_dec * dginit;
                    dstest = P.s_curv * dginit;
                    count = 0; bracketed = false; touched = false; lo = 0.0; hi = P.max_step;
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = xp[i] + stp * d[i];
                    phase = PH_LS;
                }
            }
        }

        // ---- retire ----------------------------------------------------------------------------
        if (__any_sync(FULL, finish != 0)) {
            const bool writer = !REP || LN::giw() == 0;   // replicas hold identical results: replica 0 writes them
            if (finish && writer) {
                store_x<LPT>(a.x + (size_t)prob * n, N, lig, x);
                if (lig == 0) {
                    if (a.f_out) a.f_out[prob] = fx;
                    if (a.status) a.status[prob] = ret;
                    if (a.iters) a.iters[prob] = k;
                    if (a.evals) a.evals[prob] = evals;
                    my_evals += (unsigned long long)evals;
                }
            }
            if (a.coeffs || a.T)
                emit_trajectory<S, LPT>(FULL, lig, finish ? N : 0, rounds, pv, x,
                                        (finish && writer && a.coeffs) ? a.coeffs + (size_t)prob * N * 3 * 2 * S : nullptr,
                                        (finish && writer && a.T) ? a.T + (size_t)prob * N : nullptr);
            if (finish) phase = PH_FETCH;
            __syncwarp();   // emit_trajectory read the group's head / tail; the fetch of the next trip overwrites them
        }
    }
    if (a.total_evals && my_evals) atomicAdd(a.total_evals, my_evals);
#ifdef MINCOB_TIMING
    if (threadIdx.x == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMax(a.total_evals + 2, t);
    }
#endif
}

// One launch = the lbfgs_evaluate_t callback body for every problem of the batch.
template <int S, int LPT, int THREADS>
__global__ void __launch_bounds__(THREADS) evaluate_kernel(const DevParams P, const BatchArgs a) {
    using LN = Lanes<LPT>;
    constexpr int GPB = (THREADS / 32) * LN::GPW;
    const int lig = LN::lig();
    const int N = a.N, n = N + 3 * (N - 1);
    const int groups = gridDim.x * GPB;
    const int rounds = (a.B + groups - 1) / groups;
    int p = blockIdx.x * GPB + LN::gib();
    for (int it = 0; it < rounds; ++it, p += groups) {
        const bool live = LN::real() && p < a.B;
        const int pp = live ? p : 0;
        const ProblemView pv = view_global<S>(a, pp, lig);
        double xv[4], gt, gq[3];
        load_x<LPT>(a.x_in + (size_t)pp * n, live ? N : 0, lig, xv);
        double xq[3] = {xv[1], xv[2], xv[3]};
        const double f = cost_functional<S, LPT, false>(P, 0xffffffffu, lig, live ? N : 0, N > 2 ? N - 2 : 0, pv, typename SplineReg<S, LPT>::ST_t(), xv[0], xq, gt, gq);
        if (live) {
            double gv[4] = {gt, gq[0], gq[1], gq[2]};
            store_x<LPT>(a.g_out + (size_t)p * n, N, lig, gv);
            if (lig == 0) a.f_out[p] = f;
        }
    }
}

}  // namespace mincob
