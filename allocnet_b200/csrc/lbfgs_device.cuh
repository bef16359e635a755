// ============================================================================
// lbfgs_device.cuh -- persistent batched L-BFGS driver, one LPT-lane group per trajectory.
//
// Control flow restates lbfgs::lbfgs_optimize (src/planner/include/gcopter/lbfgs.hpp:434-717)
// and line_search_lewisoverton (:276-384) of the reference: same tests, same order, same
// return codes (:135-184).  It is written as a per-group state machine with exactly ONE
// cost-functional call site per loop trip so that the 32/LPT groups of a warp, each somewhere
// else in its own line search, execute the expensive evaluation convergently.  A group that
// finishes writes its results (x, f, status, iterations, evaluations, Trajectory-order
// coefficients, durations) and pulls the next problem from a global counter, so there is no
// tail of half-empty launches and no per-iteration HBM traffic for optimizer state:
// x, g, xp, gp, d live in registers (lane i owns tau_i and q_i), the (s, y) history in
// shared memory.
// ============================================================================
#pragma once
#include "minco_device.cuh"

namespace mincob {



template <int S>
__device__ __forceinline__ ProblemView view_of(const BatchArgs &a, int p) {
    ProblemView pv;
    pv.head = a.head + (size_t)p * S * 3;
    pv.tail = a.tail + (size_t)p * S * 3;
    pv.planes = a.hpolys ? a.hpolys + (size_t)p * a.N * a.K * 4 : nullptr;
    pv.hrows = a.hrows ? a.hrows + (size_t)p * a.N : nullptr;
    pv.K = a.hpolys ? a.K : 0;
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
__device__ __noinline__ void emit_trajectory(unsigned mask, int lig, int N, const ProblemView pv,
                                             const double (&xv)[4], double *coeffs, double *Tout) {
    constexpr int D = 2 * S, b = S - 1;
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
    Spline<S, LPT> sp;
    double chat[D][3];
    spline_solve<S, LPT>(mask, lig, N, T, P0, P1, hd, td, sp, chat);
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

template <int S, int LPT, int THREADS>
__global__ void __launch_bounds__(THREADS) optimize_kernel(const DevParams P, const BatchArgs a) {
    constexpr int GPB = THREADS / LPT;
    const int lig = (threadIdx.x & 31) % LPT;
    const int gib = threadIdx.x / LPT;
    const unsigned mask = group_mask<LPT>();
    const int N = a.N, n = N + 3 * (N - 1), m = P.mem, past = P.past;

    extern __shared__ double smem[];
    const int per_group = 2 * m * 4 * LPT + 2 * m + (past > 0 ? past : 1);
    double *hs = smem + (size_t)gib * per_group;
    double *hy = hs + m * 4 * LPT;
    double *alpha = hy + m * 4 * LPT;
    double *ysv = alpha + m;
    double *pf = ysv + m;
    (void)GPB;

    int phase = PH_FETCH, prob = 0;
    double x[4] = {0, 0, 0, 0}, g[4], xp[4] = {0, 0, 0, 0}, gp[4] = {0, 0, 0, 0}, d[4] = {0, 0, 0, 0};
    double fx = 0.0, stp = 0.0, finit = 0.0, dgtest = 0.0, dstest = 0.0, lo = 0.0, hi = 0.0;
    int count = 0, k = 0, end = 0, bound = 0, evals = 0;
    bool bracketed = false, touched = false;
    unsigned long long my_evals = 0ull;

    for (;;) {
        if (phase == PH_FETCH) {
            int p = 0;
            if (lig == 0) p = atomicAdd(a.counter, 1);
            p = __shfl_sync(mask, p, 0, LPT);
            if (p >= a.B) {
                phase = PH_IDLE;
                prob = 0;
            } else {
                prob = p;
                load_x<LPT>(a.x + (size_t)p * n, N, lig, x);
                phase = PH_FIRST;
            }
        }
        __syncwarp();
        if (__all_sync(0xffffffffu, phase == PH_IDLE)) break;

        const ProblemView pv = view_of<S>(a, prob);
        double xq[3] = {x[1], x[2], x[3]}, gq[3];
        const double f = cost_functional<S, LPT>(P, mask, lig, phase == PH_IDLE ? 0 : N, pv, x[0], xq, g[0], gq);
        g[1] = gq[0]; g[2] = gq[1]; g[3] = gq[2];

        if (phase == PH_IDLE) continue;

        int finish = 0, ret = 0;     // finish: 1 = done with this problem
        bool start_ls = false;
        if (phase == PH_FIRST) {
            evals = 1;
            fx = f;
            if (lig == 0) pf[0] = fx;
            __syncwarp(mask);
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = -g[i];
            const double gn = ginf<LPT>(mask, g), xn = ginf<LPT>(mask, x);
            k = 0;
            if (gn / fmax(1.0, xn) < P.g_eps) {
                ret = LBFGS_CONVERGENCE;
                finish = 1;
            } else {
                stp = 1.0 / sqrt(gdot<LPT>(mask, d, d));
                k = 1; end = 0; bound = 0;
                start_ls = true;
            }
        } else {  // PH_LS: one trial point of line_search_lewisoverton evaluated
            ++count; ++evals;
            fx = f;
            int fail = 0;
            bool done = false;
            if (isinf(f) || isnan(f)) {
                fail = LBFGSERR_INVALID_FUNCVAL;
            } else if (f > finit + stp * dgtest) {
                hi = stp; bracketed = true;
            } else if (gdot<LPT>(mask, g, d) < dstest) {
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
            if (fail) {  // lbfgs.hpp:570-577: revert to the last good iterate
#pragma unroll
                for (int i = 0; i < 4; ++i) { x[i] = xp[i]; g[i] = gp[i]; }
                ret = fail;
                finish = 1;
            } else if (done) {
                const double gn = ginf<LPT>(mask, g), xn = ginf<LPT>(mask, x);
                if (gn / fmax(1.0, xn) < P.g_eps) {
                    ret = LBFGS_CONVERGENCE; finish = 1;
                }
                if (!finish && past > 0) {
                    if (past <= k) {
                        const double rate = fabs(pf[k % past] - fx) / fmax(1.0, fabs(fx));
                        if (rate < P.delta) { ret = LBFGS_STOP; finish = 1; }
                    }
                    if (!finish) {
                        __syncwarp(mask);
                        if (lig == 0) pf[k % past] = fx;
                        __syncwarp(mask);
                    }
                }
                if (!finish && P.max_iter != 0 && P.max_iter <= k) { ret = LBFGSERR_MAXIMUMITERATION; finish = 1; }
                if (!finish) {
                    ++k;
                    double sv[4], yv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        sv[i] = x[i] - xp[i]; yv[i] = g[i] - gp[i];
                        hs[(end * 4 + i) * LPT + lig] = sv[i];
                        hy[(end * 4 + i) * LPT + lig] = yv[i];
                    }
                    const double ys = gdot<LPT>(mask, yv, sv), yy = gdot<LPT>(mask, yv, yv);
                    if (lig == 0) ysv[end] = ys;
                    __syncwarp(mask);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] = -g[i];
                    const double cau = gdot<LPT>(mask, sv, sv) * sqrt(gdot<LPT>(mask, gp, gp)) * P.cautious;
                    if (ys > cau) {
                        ++bound;
                        bound = m < bound ? m : bound;
                        end = (end + 1) % m;
                        int j = end;
                        for (int i = 0; i < bound; ++i) {
                            j = (j + m - 1) % m;
                            double sj[4], yj[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) { sj[u] = hs[(j * 4 + u) * LPT + lig]; yj[u] = hy[(j * 4 + u) * LPT + lig]; }
                            const double aj = gdot<LPT>(mask, sj, d) / ysv[j];
                            if (lig == 0) alpha[j] = aj;
#pragma unroll
                            for (int u = 0; u < 4; ++u) d[u] += (-aj) * yj[u];
                        }
                        __syncwarp(mask);
                        const double sc0 = ys / yy;
#pragma unroll
                        for (int u = 0; u < 4; ++u) d[u] *= sc0;
                        for (int i = 0; i < bound; ++i) {
                            double sj[4], yj[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) { sj[u] = hs[(j * 4 + u) * LPT + lig]; yj[u] = hy[(j * 4 + u) * LPT + lig]; }
                            const double beta = gdot<LPT>(mask, yj, d) / ysv[j];
                            const double cf = alpha[j] - beta;
#pragma unroll
                            for (int u = 0; u < 4; ++u) d[u] += cf * sj[u];
                            j = (j + 1) % m;
                        }
                    }
                    stp = 1.0;
                    start_ls = true;
                }
            }
        }
        if (start_ls) {  // entry of line_search_lewisoverton (lbfgs.hpp:276-310)
#pragma unroll
            for (int i = 0; i < 4; ++i) { xp[i] = x[i]; gp[i] = g[i]; }
            int fail = 0;
            double dginit = 0.0;
            if (!(stp > 0.0)) fail = LBFGSERR_INVALIDPARAMETERS;
            else {
                dginit = gdot<LPT>(mask, gp, d);
                if (0.0 < dginit) fail = LBFGSERR_INCREASEGRADIENT;
            }
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
        if (finish) {
            store_x<LPT>(a.x + (size_t)prob * n, N, lig, x);
            if (lig == 0) {
                if (a.f_out) a.f_out[prob] = fx;
                if (a.status) a.status[prob] = ret;
                if (a.iters) a.iters[prob] = k;
                if (a.evals) a.evals[prob] = evals;
                my_evals += (unsigned long long)evals;
            }
            if (a.coeffs || a.T)
                emit_trajectory<S, LPT>(mask, lig, N, pv, x, a.coeffs ? a.coeffs + (size_t)prob * N * 3 * 2 * S : nullptr,
                                        a.T ? a.T + (size_t)prob * N : nullptr);
            phase = PH_FETCH;
        }
    }
    if (a.total_evals && my_evals) atomicAdd(a.total_evals, my_evals);
}

// One launch = the lbfgs_evaluate_t callback body for every problem of the batch.
template <int S, int LPT, int THREADS>
__global__ void __launch_bounds__(THREADS) evaluate_kernel(const DevParams P, const BatchArgs a) {
    const int lig = (threadIdx.x & 31) % LPT;
    const unsigned mask = group_mask<LPT>();
    const int N = a.N, n = N + 3 * (N - 1);
    const int groups = gridDim.x * (THREADS / LPT);
    const int rounds = (a.B + groups - 1) / groups;
    int p = blockIdx.x * (THREADS / LPT) + threadIdx.x / LPT;
    for (int it = 0; it < rounds; ++it, p += groups) {
        const bool live = p < a.B;
        const int pp = live ? p : 0;
        const ProblemView pv = view_of<S>(a, pp);
        double xv[4], gt, gq[3];
        load_x<LPT>(a.x_in + (size_t)pp * n, live ? N : 0, lig, xv);
        double xq[3] = {xv[1], xv[2], xv[3]};
        const double f = cost_functional<S, LPT>(P, mask, lig, live ? N : 0, pv, xv[0], xq, gt, gq);
        if (live) {
            double gv[4] = {gt, gq[0], gq[1], gq[2]};
            store_x<LPT>(a.g_out + (size_t)p * n, N, lig, gv);
            if (lig == 0) a.f_out[p] = f;
        }
    }
}

}  // namespace mincob
