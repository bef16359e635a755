// ============================================================================
// kernels_inst.cu -- one (S, LPT) instantiation of every kernel plus its launchers.
// Built once per pair with -DMINCOB_S=.. -DMINCOB_LPT=.. (allocnet_b200/build.py) so the six
// objects compile in parallel.
// ============================================================================
#include "../../include/mincob.h"
#include <cstdlib>

#include "launch.h"
#include "lbfgs_device.cuh"

#ifndef MINCOB_S
#error "compile with -DMINCOB_S=3|4 -DMINCOB_LPT=8|16|32"
#endif

namespace mincob {

// ---- MINCO building-block kernels (setParameters/getEnergy/.../propogateGrad) -------------

template <int S, int LPT, int THREADS>
__global__ void __launch_bounds__(THREADS) minco_kernel(const MincoArgs a, int propagate) {
    constexpr int D = 2 * S, b = S - 1;
    using LN = Lanes<LPT>;
    constexpr int GPB = (THREADS / 32) * LN::GPW;
    const int lig = LN::lig();
    const unsigned mask = 0xffffffffu;  // uniform control flow: whole warp participates in every shuffle
    const int N = a.N;
    const int groups = gridDim.x * GPB;
    const int rounds = (a.B + groups - 1) / groups;
    int p = blockIdx.x * GPB + LN::gib();
    for (int it = 0; it < rounds; ++it, p += groups) {
        const bool live = LN::real() && p < a.B;
        const int pp = live ? p : 0;
        const int Ne = live ? N : 0;
        const bool active = lig < Ne;
        const double *head = a.head + (size_t)pp * S * 3, *tail = a.tail + (size_t)pp * S * 3;
        const double *q = a.inPs + (size_t)pp * (N - 1) * 3;
        double P0[3], P1[3], hd[b][3], td[b][3];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            P0[x] = (lig == 0) ? head[x] : ((active && lig >= 1) ? q[(lig - 1) * 3 + x] : 0.0);
            P1[x] = (lig == Ne - 1) ? tail[x] : ((active) ? q[lig * 3 + x] : 0.0);
        }
#pragma unroll
        for (int d = 0; d < b; ++d)
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                hd[d][x] = (lig == 0) ? head[(d + 1) * 3 + x] : 0.0;
                td[d][x] = (lig == Ne - 1) ? tail[(d + 1) * 3 + x] : 0.0;
            }
        const double T = active ? a.ts[(size_t)pp * N + lig] : 1.0;
        SplineReg<S, LPT> sp;
        double chat[D][3];
        spline_solve<S, LPT>(mask, lig, Ne, N > 2 ? N - 2 : 0, T, P0, P1, hd, td, sp, chat);
        if (!propagate) {
            double e, G[D][3], gT;
            energy_partials<S, LPT>(sp, chat, active, e, G, gT);
            e = group_sum<LPT>(mask, e);
            if (active) {
#pragma unroll
                for (int k = 0; k < D; ++k)
#pragma unroll
                    for (int x = 0; x < 3; ++x) {
                        const size_t row = (size_t)p * D * N + (size_t)D * lig + k;
                        if (a.coeffs_asc) a.coeffs_asc[row * 3 + x] = sp.c[k][x];
                        if (a.gdC) a.gdC[row * 3 + x] = G[k][x];
                        if (a.flat) a.flat[(((size_t)p * N + lig) * 3 + x) * D + k] = sp.c[D - 1 - k][x];
                    }
                if (a.gdT) a.gdT[(size_t)p * N + lig] = gT;
                if (a.energy && lig == 0) a.energy[p] = e;
            }
        } else {
            double G[D][3], gq[3], gT;
#pragma unroll
            for (int k = 0; k < D; ++k)
#pragma unroll
                for (int x = 0; x < 3; ++x)
                    G[k][x] = active ? a.gdC_in[((size_t)pp * D * N + (size_t)D * lig + k) * 3 + x] : 0.0;
            const double gTp = active ? a.gdT_in[(size_t)pp * N + lig] : 0.0;
            spline_adjoint<S, LPT>(mask, lig, Ne, N > 2 ? N - 2 : 0, sp, G, gTp, gq, gT);
            if (active) {
                a.gradByTimes[(size_t)p * N + lig] = gT;
                if (lig >= 1) {
#pragma unroll
                    for (int x = 0; x < 3; ++x) a.gradByPoints[((size_t)p * (N - 1) + lig - 1) * 3 + x] = gq[x];
                }
            }
        }
    }
}

// ---- feasibility report: one lane per piece, `samples`+1 points per piece, group-wide max ---------------
template <int S, int LPT, int THREADS>
__global__ void __launch_bounds__(THREADS) check_kernel(const CheckArgs a) {
    constexpr int D = 2 * S;
    using LN = Lanes<LPT>;
    constexpr int GPB = (THREADS / 32) * LN::GPW;
    const int lig = LN::lig();
    const unsigned mask = 0xffffffffu;
    const int N = a.N;
    const int groups = gridDim.x * GPB;
    const int rounds = (a.B + groups - 1) / groups;
    int p = blockIdx.x * GPB + LN::gib();
    for (int it = 0; it < rounds; ++it, p += groups) {
        const bool live = LN::real() && p < a.B && lig < N;
        double vmax = 0.0, amax = 0.0, jmax = 0.0, cmax = -1.0e300;
        if (live) {
            const double *c = a.coeffs + ((size_t)p * N + lig) * 3 * D;   // [3][2S], k = 0 highest power
            const double T = a.T[(size_t)p * N + lig];
            const bool planes = a.hpolys && a.hrows && a.K > 0;
            const double *hp = planes ? a.hpolys + ((size_t)p * N + lig) * a.K * 4 : nullptr;
            const int rows = planes ? min(a.hrows[(size_t)p * N + lig], a.K) : 0;
            for (int j = 0; j <= a.samples; ++j) {
                const double t = T * j / a.samples;
                double pos[3], v2 = 0.0, a2 = 0.0, j2 = 0.0;
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    // Horner on the value and its first three derivatives (scaled by 1/d!) together;
                    // descending coefficients as in Piece<D>, trajectory.hpp:75-133
                    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        q3 = fma(q3, t, q2);
                        q2 = fma(q2, t, q1);
                        q1 = fma(q1, t, q0);
                        q0 = fma(q0, t, c[x * D + k]);
                    }
                    q2 *= 2.0; q3 *= 6.0;
                    pos[x] = q0; v2 += q1 * q1; a2 += q2 * q2; j2 += q3 * q3;
                }
                vmax = fmax(vmax, v2); amax = fmax(amax, a2); jmax = fmax(jmax, j2);
                for (int k = 0; k < rows; ++k) {
                    const Plane h = load_plane<false>(hp + 4 * k);
                    cmax = fmax(cmax, fma(h.x, pos[0], fma(h.y, pos[1], fma(h.z, pos[2], h.w))));
                }
            }
        }
        vmax = group_max<LPT>(mask, vmax); amax = group_max<LPT>(mask, amax);
        jmax = group_max<LPT>(mask, jmax); cmax = group_max<LPT>(mask, cmax);
        if (LN::real() && p < a.B && lig == 0) {
            double *o = a.out + (size_t)p * 4;
            o[0] = sqrt(vmax); o[1] = sqrt(amax); o[2] = sqrt(jmax); o[3] = cmax;
        }
    }
}

// ---- exact rate maxima: one lane per piece ----------------------------------------------------------------
// q_d(t) = |p^(d)(t)|^2 on [0, T] has its maximum at an end point or where q_d' = 2 p^(d).p^(d+1) vanishes.  The
// reference forms q_d' in monomial form and isolates its roots with Sturm sequences (root_finder.hpp, tolerance
// FLT_EPSILON / T); here every lane walks `grid` sub-intervals of its piece, keeps the largest q_d seen, and
// wherever p^(d).p^(d+1) changes sign bisects the bracket to the last bit and takes q_d there.  Two stationary
// points inside one sub-interval (1/grid of a piece) are seen only through the grid values; with grid = 128 that
// is far below the reference's own tolerance for trajectories an optimizer produces.
template <int D>
__device__ __forceinline__ void deriv3(const double *c, double t, int d, double (&r)[3]) {
    // d-th derivative of the three axis polynomials; c[x*D + k] multiplies t^(D-1-k) (trajectory.hpp:75-133)
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        double acc = 0.0;
        for (int k = 0; k < D - d; ++k) {
            double f = 1.0;
            for (int u = 0; u < d; ++u) f *= (double)(D - 1 - k - u);
            acc = fma(acc, t, c[x * D + k] * f);
        }
        r[x] = acc;
    }
}
template <int S, int LPT, int THREADS>
__global__ void __launch_bounds__(THREADS) maxrate_kernel(const RateArgs a) {
    constexpr int D = 2 * S;
    using LN = Lanes<LPT>;
    constexpr int GPB = (THREADS / 32) * LN::GPW;
    const int lig = LN::lig();
    const unsigned mask = 0xffffffffu;
    const int N = a.N;
    const int groups = gridDim.x * GPB;
    const int rounds = (a.B + groups - 1) / groups;
    int p = blockIdx.x * GPB + LN::gib();
    for (int it = 0; it < rounds; ++it, p += groups) {
        const bool live = LN::real() && p < a.B && lig < N;
        double best[3] = {0.0, 0.0, 0.0};
        if (live) {
            double c[3 * D];
#pragma unroll
            for (int i = 0; i < 3 * D; ++i) c[i] = a.coeffs[((size_t)p * N + lig) * 3 * D + i];
            const double T = a.T[(size_t)p * N + lig];
            for (int d = 1; d <= 3; ++d) {
                double bq = 0.0, tp = 0.0, gp = 0.0;
                for (int i = 0; i <= a.grid; ++i) {
                    const double t = (i == a.grid) ? T : T * i / a.grid;
                    double r[3], rn[3];
                    deriv3<D>(c, t, d, r);
                    deriv3<D>(c, t, d + 1, rn);
                    const double g = r[0] * rn[0] + r[1] * rn[1] + r[2] * rn[2];
                    bq = fmax(bq, r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                    if (i > 0 && ((gp < 0.0) != (g < 0.0)) && gp != 0.0 && g != 0.0) {
                        double lo = tp, hi = t;
                        const bool lo_neg = gp < 0.0;
                        for (int b = 0; b < 64; ++b) {
                            const double mid = 0.5 * (lo + hi);
                            if (!(lo < mid && mid < hi)) break;
                            deriv3<D>(c, mid, d, r);
                            deriv3<D>(c, mid, d + 1, rn);
                            const double gm = r[0] * rn[0] + r[1] * rn[1] + r[2] * rn[2];
                            if ((gm < 0.0) == lo_neg && gm != 0.0) lo = mid; else hi = mid;
                        }
                        deriv3<D>(c, 0.5 * (lo + hi), d, r);
                        bq = fmax(bq, r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                    }
                    tp = t; gp = g;
                }
                best[d - 1] = bq;
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) best[d] = group_max<LPT>(mask, best[d]);
        if (LN::real() && p < a.B && lig == 0) {
            double *o = a.out + (size_t)p * 3;
            o[0] = sqrt(best[0]); o[1] = sqrt(best[1]); o[2] = sqrt(best[2]);
        }
    }
}

}  // namespace mincob

namespace {
using namespace mincob;
#ifndef MINCOB_THREADS
#define MINCOB_THREADS 128   // threads per block of every kernel here (experiments: 256 / 384 with MINCOB_LOCKSTEP)
#endif
constexpr int S = MINCOB_S, LPT = MINCOB_LPT, THREADS = MINCOB_THREADS;
constexpr int WARPS = THREADS / 32;
constexpr int GPB = WARPS * Lanes<LPT>::GPW;   // trajectories per block (throughput mapping)
constexpr int SPB = WARPS * Lanes<LPT>::SPW;   // group slots per block (GPB + one dummy group per warp when LPT does not divide 32)

LaunchResult ok(cudaError_t e) { return LaunchResult{e, 0, 0, 0}; }

LaunchResult launch_evaluate(cudaStream_t st, int sm_count, const DevParams &dp, const BatchArgs &a) {
    int blocks = (a.B + GPB - 1) / GPB;
    const int cap = sm_count * 16;
    if (blocks > cap) blocks = cap;
    evaluate_kernel<S, LPT, THREADS><<<blocks, THREADS, 0, st>>>(dp, a);
    return ok(cudaGetLastError());
}

// Shared memory per block and grid of the persistent optimize kernel.  Half-planes are staged in
// shared memory when at least MINCOB_MINB blocks per SM still fit; otherwise they are read from global.
struct OptPlan {
    int psmem, rep, blocks;   // psmem: 0 rows read from global memory, 1 staged in shared memory
    size_t smem, hist_bytes, mult_bytes, park_bytes;
    int code;
    cudaError_t err;
};
// history depth the register-resident two-loop recursion is compiled for (upstream default of this build, params.py)
#ifndef MINCOB_FASTMEM
#define MINCOB_FASTMEM 8
#endif
constexpr int FASTMEM = MINCOB_FASTMEM;   // 0: every depth takes the rolled loops (experiments)
using OptKernel = void (*)(const DevParams, const BatchArgs);
template <int PSM, int MEM, bool FRZ>
static OptKernel opt_kernel(bool rep) {
    if (rep) return optimize_kernel<S, LPT, THREADS, PSM, MEM, true, FRZ>;
    return optimize_kernel<S, LPT, THREADS, PSM, MEM, false, FRZ>;
}
template <int MEM, bool FRZ>
static OptKernel opt_kernel_m(int psm, bool rep) {
    return psm == 1 ? opt_kernel<1, MEM, FRZ>(rep) : opt_kernel<0, MEM, FRZ>(rep);
}
// frz: the fixed-time specialisation exists for the default history depth; other depths run the generic kernel, which
// honours DevParams::freeze at run time (same bits)
static OptKernel pick_kernel(int psm, int mem, bool rep, bool frz) {
    if (mem == FASTMEM) return (frz && FASTMEM > 0) ? opt_kernel_m<FASTMEM, true>(psm, rep) : opt_kernel_m<FASTMEM, false>(psm, rep);
    return opt_kernel_m<0, false>(psm, rep);
}
static int blocks_per_sm(OptKernel kern, size_t smem, cudaError_t &e) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { e = cudaSuccess; (void)cudaGetLastError(); return 0; }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    return e == cudaSuccess ? per_sm : 0;
}
// mapping: MINCOB_MAP_AUTO picks the latency mapping (one warp per trajectory) when the batch cannot give every
// resident warp a trajectory of its own anyway; the two mappings produce the same bits (penalty_piece), so the choice
// changes timing only.
static size_t smem_bytes(int N, int K, const DevParams &dp, int psmem) {
    return ((size_t)GPB * optimize_group_doubles(S, N, K, dp.mem, dp.past, psmem, LPT) +
            (size_t)(SPB - GPB) * optimize_small_doubles(S, dp.mem, dp.past, LPT)) * sizeof(double);
}
// MINCOB_NO_FRZ in the environment sends the fixed-time mode through the generic kernel (DevParams::freeze at run time):
// the parity suite compares the two, they must agree bit for bit.
static bool use_frz(const DevParams &dp) { return dp.freeze != 0 && getenv("MINCOB_NO_FRZ") == nullptr; }
static OptPlan plan_optimize(int sm_count, const DevParams &dp, const BatchArgs &a) {
    OptPlan pl{0, 0, 0, 0, 0, 0, 0, 0, cudaSuccess};
    const bool frz = use_frz(dp);
    const bool have = a.hpolys && a.hrows && a.K > 0;
    const bool can_rep = Lanes<LPT>::GPW > 1;
    int per_sm = 0;
    // occupancy does not depend on REP (same registers cap, same shared memory): plan with the throughput kernel
    if (have && dp.penalties) {
        pl.smem = smem_bytes(a.N, a.K, dp, 1);
        per_sm = blocks_per_sm(pick_kernel(1, dp.mem, false, frz), pl.smem, pl.err);
        if (pl.err != cudaSuccess) return pl;
        pl.psmem = per_sm >= MINCOB_MINB || per_sm * WARPS >= 8;   // staging must leave at least 8 warps per SM resident
#ifdef MINCOB_GLOBAL_PLANES   // experiment: never stage half-planes in shared memory (occupancy then depends on registers only)
        pl.psmem = 0;
#endif
    }
    if (!pl.psmem) {
        pl.smem = smem_bytes(a.N, a.K, dp, 0);
        per_sm = blocks_per_sm(pick_kernel(false, dp.mem, false, frz), pl.smem, pl.err);
        if (pl.err != cudaSuccess) return pl;
    }
    if (per_sm < 1) { pl.code = MINCOB_E_INVALID; return pl; }
    pl.blocks = per_sm * sm_count;
    // (the latency mapping keeps one flag bit per penalty sample: kappa <= 31, else the throughput mapping runs)
    pl.rep = can_rep && dp.kappa <= 31 && (dp.mapping == MINCOB_MAP_LATENCY || (dp.mapping == MINCOB_MAP_AUTO && a.B <= pl.blocks * WARPS));
    if (pl.rep) {
        per_sm = blocks_per_sm(pick_kernel(pl.psmem, dp.mem, true, frz), pl.smem, pl.err);
        if (pl.err != cudaSuccess) return pl;
        if (per_sm < 1) { pl.code = MINCOB_E_INVALID; return pl; }
        pl.blocks = per_sm * sm_count;
    }
    const int per_block = pl.rep ? WARPS : GPB;
    const int need = (a.B + per_block - 1) / per_block;
    if (pl.blocks > need) pl.blocks = need;
    pl.hist_bytes = (size_t)pl.blocks * SPB * dp.mem * LPT * 8 * sizeof(double);
    pl.hist_bytes = (pl.hist_bytes + 255) & ~(size_t)255;
    pl.mult_bytes = (size_t)pl.blocks * SplineReg<S, LPT>::NM * THREADS * sizeof(double);
    pl.park_bytes = (size_t)pl.blocks * 12 * THREADS * sizeof(double);
    return pl;
}

// one allocation: [history slabs | multiplier slabs | parked state]; sized for the larger of the two mappings
size_t optimize_scratch(int sm_count, const DevParams &dp, const BatchArgs &a) {
    const OptPlan pl = plan_optimize(sm_count, dp, a);
    return pl.hist_bytes + pl.mult_bytes + pl.park_bytes;
}

LaunchResult launch_optimize(cudaStream_t st, int sm_count, const DevParams &dp, const BatchArgs &a) {
    const OptPlan pl = plan_optimize(sm_count, dp, a);
    if (pl.err != cudaSuccess) return ok(pl.err);
    if (pl.code) return LaunchResult{cudaSuccess, pl.code, pl.smem, 0};
    cudaError_t e;
    if ((e = cudaMemsetAsync(a.counter, 0, sizeof(int), st)) != cudaSuccess) return ok(e);
    if ((e = cudaMemsetAsync(a.total_evals, 0, 128 * sizeof(unsigned long long), st)) != cudaSuccess) return ok(e);
#ifdef MINCOB_TIMING
    cudaMemsetAsync(a.total_evals + 1, 0xff, sizeof(unsigned long long), st);
    cudaMemsetAsync(a.total_evals + 3, 0xff, sizeof(unsigned long long), st);
#endif
    BatchArgs b = a;
    b.mult = reinterpret_cast<double *>(reinterpret_cast<char *>(a.hist) + pl.hist_bytes);
    b.lpark = reinterpret_cast<double *>(reinterpret_cast<char *>(a.hist) + pl.hist_bytes + pl.mult_bytes);
    const bool frz = use_frz(dp);
    pick_kernel(pl.psmem, dp.mem, pl.rep, frz)<<<pl.blocks, THREADS, pl.smem, st>>>(dp, b);
    LaunchResult r = ok(cudaGetLastError());
    r.mapping = pl.rep ? MINCOB_MAP_LATENCY : MINCOB_MAP_THROUGHPUT;
    return r;
}

LaunchResult launch_minco(cudaStream_t st, int sm_count, const MincoArgs &a, int propagate) {
    int blocks = (a.B + GPB - 1) / GPB;
    const int cap = sm_count * 16;
    if (blocks > cap) blocks = cap;
    minco_kernel<S, LPT, THREADS><<<blocks, THREADS, 0, st>>>(a, propagate);
    return ok(cudaGetLastError());
}

LaunchResult launch_check(cudaStream_t st, int sm_count, const CheckArgs &a) {
    int blocks = (a.B + GPB - 1) / GPB;
    const int cap = sm_count * 16;
    if (blocks > cap) blocks = cap;
    check_kernel<S, LPT, THREADS><<<blocks, THREADS, 0, st>>>(a);
    return ok(cudaGetLastError());
}

LaunchResult launch_maxrates(cudaStream_t st, int sm_count, const RateArgs &a) {
    int blocks = (a.B + GPB - 1) / GPB;
    const int cap = sm_count * 16;
    if (blocks > cap) blocks = cap;
    maxrate_kernel<S, LPT, THREADS><<<blocks, THREADS, 0, st>>>(a);
    return ok(cudaGetLastError());
}

const LaunchTable kTable = {launch_evaluate, launch_optimize, launch_minco, optimize_scratch, launch_check, launch_maxrates};
}  // namespace

#define MINCOB_CAT_(a, b, c) mincob_table_##a##_##b
#define MINCOB_CAT(a, b) MINCOB_CAT_(a, b, )
const mincob::LaunchTable *MINCOB_CAT(MINCOB_S, MINCOB_LPT)() { return &kTable; }
