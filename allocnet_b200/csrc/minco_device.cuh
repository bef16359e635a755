// ============================================================================
// minco_device.cuh -- device-side MINCO cost functional for sm_100a.
//
// One LANE per trajectory piece, LPT (8/16/32) lanes per trajectory, 32/LPT trajectories per
// warp; all cross-piece traffic is warp shuffles inside the LPT-lane group, so groups of one
// warp are independent (every shuffle names only its own group's lanes).
//
// What is computed (same call sequence as the upstream GCOPTER costFunctional restated in
// SURVEY.md Appendix B.1; NOT in /root/reference):
//   forwardT -> MINCO setParameters (banded solve, Appendix A.2) -> getEnergy +
//   getEnergyPartialGradBy{Coeffs,Times} (A.3) -> attachPenaltyFunctional (B.2, smoothedL1 of
//   gcopter/firi.hpp:60-84) -> propogateGrad (A.4) -> + rho*sum(T) -> backwardGradT.
//
// How the banded solve is done here.  The 2S*N x 2S*N band of minco.hpp is a 48-step serial
// elimination for N=8.  Its rows say "the spline is C^(2S-2) at the waypoints", which is the
// stationarity of E = sum_i int |p_i^(S)|^2 with respect to the junction derivatives
// y_j = (p',..,p^(S-1))(t_j).  In Hermite form c_i = H(T_i) s_i, s_i = [start state; end state],
//   E = sum_i s_i^T W(T_i) s_i,  W(T) = T^(1-2S) L What L,  L = diag(1,T,..,T^(S-1)) twice,
// so y solves a symmetric positive definite BLOCK-TRIDIAGONAL system with (S-1)x(S-1) blocks
// and one block row per junction.  Lane j builds block row j from T_{j-1}, T_j and the system is
// solved by block elimination that sweeps from lane to lane with shuffles (N-2 short rounds of
// (S-1)x(S-1) algebra instead of 2S*N scalar pivots); the sweep is a ROLLED loop, so the whole solve
// is ~150 instructions of code -- what matters more than its latency, because the persistent
// optimize kernel is instruction-cache bound (profiles/).  The multipliers are kept so the adjoint
// solve (the matrix is symmetric) only touches right-hand sides.  oracle/reduced_proto.py is the readable numpy statement of the same algebra and
// tests/test_reduced_formulation.py checks it against the banded oracle.
// ============================================================================
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "args.h"
#include "hermite_constants.cuh"

#ifndef MINCOB_UNROLL_JJ
#define MINCOB_UNROLL_JJ 1   // samples per trip of the rolled phase-2 loop of penalty_piece
#endif
#ifndef MINCOB_JB
#define MINCOB_JB 6          // samples whose positions phase 1 of penalty_piece holds in registers at once
#endif
#ifndef MINCOB_PLANE_PREFETCH
#define MINCOB_PLANE_PREFETCH 1   // +1 % (N = 8) ... +3 % (N = 5): profiles/r02_ab.log
#endif
#ifndef MINCOB_UNROLL_KG
#define MINCOB_UNROLL_KG 2   // same, when the rows are read from global memory (polytopes too large to stage)
#endif
#ifndef MINCOB_DEFER_BODY
#define MINCOB_DEFER_BODY 0  // throughput mapping: accumulate the active samples after all tests, every lane its own (experiment)
#endif
#ifndef MINCOB_UNROLL_K
#define MINCOB_UNROLL_K 1    // half-plane rows per trip of the phase-1 loop (2: -4 %, 4: -4 %, 8: -10 %; the kernel is code-size sensitive)
#endif

namespace mincob {


// d! and k!/(k-d)!: loop form (no recursion) so that they fold to literals once the surrounding
// loops are unrolled, and stay cheap straight-line code if they are not.
__host__ __device__ __forceinline__ constexpr double cfact(int d) {
    double f = 1.0;
    for (int u = 2; u <= d; ++u) f *= u;
    return f;
}
__host__ __device__ __forceinline__ constexpr double cfall(int k, int d) {
    double f = 1.0;
    for (int u = 0; u < d; ++u) f *= (k - u);
    return f;
}

// ---- lane groups -----------------------------------------------------------------------------------------
// A warp holds GPW = 32 / LPT groups of LPT consecutive lanes.  LPT a power of two (8, 16, 32): the groups tile the
// warp and the width-limited shuffles address them.  Otherwise (LPT = 5: the reference's ModelMaxSeg = 5 pieces,
// learning_planner.hpp:33 -- six trajectories per warp instead of four) the lanes past GPW * LPT form a short
// dummy group that never owns a problem, and every shuffle names its source lane explicitly.
template <int LPT>
struct Lanes {
    static constexpr bool POW2 = (LPT & (LPT - 1)) == 0;
    static constexpr int GPW = 32 / LPT;                    // real groups per warp
    static constexpr int SPW = GPW + (GPW * LPT < 32);      // group slots per warp (+ the dummy group)
    static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
    static __device__ __forceinline__ int giw() { return lane() / LPT; }   // group in warp (== GPW: the dummy group)
    static __device__ __forceinline__ int lig() { return POW2 ? (lane() & (LPT - 1)) : lane() - giw() * LPT; }
    static __device__ __forceinline__ int base() { return lane() - lig(); }
    static __device__ __forceinline__ bool real() { return POW2 || giw() < GPW; }
    // slot of this thread's group inside its block (distinct also for the dummy groups) / real group index
    static __device__ __forceinline__ int slot() { return (threadIdx.x >> 5) * SPW + giw(); }
    static __device__ __forceinline__ int gib() { return (threadIdx.x >> 5) * GPW + giw(); }
};
template <int LPT> __device__ __forceinline__ double sh_up(unsigned m, double v, int d) {
    if constexpr (Lanes<LPT>::POW2) return __shfl_up_sync(m, v, d, LPT);
    else return __shfl_sync(m, v, Lanes<LPT>::lig() >= d ? Lanes<LPT>::lane() - d : Lanes<LPT>::lane());
}
template <int LPT> __device__ __forceinline__ double sh_dn(unsigned m, double v, int d) {
    if constexpr (Lanes<LPT>::POW2) return __shfl_down_sync(m, v, d, LPT);
    else {
        const int l = Lanes<LPT>::lane();
        return __shfl_sync(m, v, (Lanes<LPT>::lig() + d < LPT && l + d < 32) ? l + d : l);
    }
}
template <int LPT> __device__ __forceinline__ double sh_xor(unsigned m, double v, int d) { return __shfl_xor_sync(m, v, d, LPT); }
// value of the group's first lane, on every lane of the group
template <int LPT, class T> __device__ __forceinline__ T group_first(unsigned m, T v) {
    if constexpr (Lanes<LPT>::POW2) return __shfl_sync(m, v, 0, LPT);
    else return __shfl_sync(m, v, Lanes<LPT>::base());
}
template <int LPT>
__device__ __forceinline__ double group_sum(unsigned m, double v) {
    if constexpr (Lanes<LPT>::POW2) {
#pragma unroll
        for (int o = LPT / 2; o > 0; o >>= 1) v += sh_xor<LPT>(m, v, o);
        return v;  // butterfly: bitwise identical on every lane of the group
    } else {
        // every lane adds the group's LPT values in lane order: bitwise identical on every lane of the group
        const int b0 = Lanes<LPT>::base();
        double s = __shfl_sync(m, v, b0);
#pragma unroll
        for (int i = 1; i < LPT; ++i) s += __shfl_sync(m, v, b0 + i);
        return s;
    }
}
template <int LPT>
__device__ __forceinline__ double group_max(unsigned m, double v) {
    if constexpr (Lanes<LPT>::POW2) {
#pragma unroll
        for (int o = LPT / 2; o > 0; o >>= 1) v = fmax(v, sh_xor<LPT>(m, v, o));
        return v;
    } else {
        const int b0 = Lanes<LPT>::base();
        double s = __shfl_sync(m, v, b0);
#pragma unroll
        for (int i = 1; i < LPT; ++i) s = fmax(s, __shfl_sync(m, v, b0 + i));
        return s;
    }
}
// Latency mapping ("one warp per trajectory"): the GPW groups of a warp hold the SAME trajectory and split the tests of
// the penalty samples among them; this ORs a flag word over the replicas (same lane-in-group), result on every replica.
template <int LPT>
__device__ __forceinline__ unsigned replica_or(unsigned m, unsigned v) {
    if constexpr (Lanes<LPT>::POW2) {
#pragma unroll
        for (int o = LPT; o < 32; o <<= 1) v |= __shfl_xor_sync(m, v, o);
        return v;
    } else {
        const int l0 = Lanes<LPT>::lig();
        unsigned s = 0u;
#pragma unroll
        for (int r = 0; r < Lanes<LPT>::GPW; ++r) s |= __shfl_sync(m, v, r * LPT + l0);
        return s;   // (the lanes of the dummy group read some real lane's words; they never own a piece)
    }
}

// tau <-> T (upstream gcopter.hpp forwardT / backwardGradT)
__device__ __forceinline__ double forward_t(double tau) {
    return tau > 0.0 ? ((0.5 * tau + 1.0) * tau + 1.0) : 1.0 / ((0.5 * tau - 1.0) * tau + 1.0);
}
__device__ __forceinline__ double backward_grad_t(double tau, double gradT) {
    if (tau > 0.0) return gradT * (tau + 1.0);
    const double den = (0.5 * tau - 1.0) * tau + 1.0;
    return gradT * (1.0 - tau) / (den * den);
}
// smoothedL1, gcopter/firi.hpp:60-84; caller guarantees x > 0.
// imu = 1/mu (hoisted: an fp64 division is ~30 instructions).
// Every sum of a product is written as an explicit fma or __dadd_rn here and in the penalty code below (plain `*` only
// where the product feeds another product or the multiplicand of an fma): the compiler is then left no choice of
// contraction, so that the copies of this code in different instantiations (the two mappings, the fixed-time kernel)
// produce the same bits.
__device__ __forceinline__ void smoothed_l1_pos(double mu, double imu, double x, double &f, double &df) {
    if (x > mu) { f = fma(-0.5, mu, x); df = 1.0; return; }
    const double r = (x * imu), r2 = (r * r), h = fma(-0.5, x, mu);
    f = ((h * r2) * r);
    df = (r2 * fma((3.0 * h), imu, (-0.5 * r)));
}

// ---- small dense helpers (b = S-1 = 2 or 3) ---------------------------------------------
template <int b>
__device__ __forceinline__ void inv_small(const double (&d)[b][b], double (&o)[b][b]);
template <>
__device__ __forceinline__ void inv_small<2>(const double (&d)[2][2], double (&o)[2][2]) {
    const double r = 1.0 / (d[0][0] * d[1][1] - d[0][1] * d[1][0]);
    o[0][0] = d[1][1] * r; o[0][1] = -d[0][1] * r;
    o[1][0] = -d[1][0] * r; o[1][1] = d[0][0] * r;
}
template <>
__device__ __forceinline__ void inv_small<3>(const double (&d)[3][3], double (&o)[3][3]) {
    const double c00 = d[1][1] * d[2][2] - d[1][2] * d[2][1];
    const double c01 = d[1][2] * d[2][0] - d[1][0] * d[2][2];
    const double c02 = d[1][0] * d[2][1] - d[1][1] * d[2][0];
    const double r = 1.0 / (d[0][0] * c00 + d[0][1] * c01 + d[0][2] * c02);
    o[0][0] = c00 * r; o[1][0] = c01 * r; o[2][0] = c02 * r;
    o[0][1] = (d[0][2] * d[2][1] - d[0][1] * d[2][2]) * r;
    o[1][1] = (d[0][0] * d[2][2] - d[0][2] * d[2][0]) * r;
    o[2][1] = (d[0][1] * d[2][0] - d[0][0] * d[2][1]) * r;
    o[0][2] = (d[0][1] * d[1][2] - d[0][2] * d[1][1]) * r;
    o[1][2] = (d[0][2] * d[1][0] - d[0][0] * d[1][2]) * r;
    o[2][2] = (d[0][0] * d[1][1] - d[0][1] * d[1][0]) * r;
}

// Where a lane keeps its block-solve multipliers between setParameters and propogateGrad: NM doubles that
// are written once and read once per evaluation.  RegStore holds them in registers (short kernels);
// GlobalStore parks them in a lane-strided global slab (coalesced, L2-resident) so that they do not
// occupy 2*NM registers across the penalty loop of the persistent optimize kernel.
template <int NM>
struct RegStore {
    double v[NM];
    __device__ __forceinline__ void put(int i, double x) { v[i] = x; }
    __device__ __forceinline__ double get(int i) const { return v[i]; }
};
struct GlobalStore {
    double *p;     // this lane's first slot
    int stride;    // distance between slots (threads sharing the slab)
    __device__ __forceinline__ void put(int i, double x) { __stcg(p + (size_t)i * stride, x); }
    __device__ __forceinline__ double get(int i) const { return __ldcg(p + (size_t)i * stride); }
};

// Per-lane (= per-piece) spline state kept between the forward solve and the adjoint.
template <int S, int LPT, class ST>
struct Spline {
    static constexpr int D = 2 * S, b = S - 1;
    static constexpr int NM = 3 * b * b;   // elimination multiplier M, inverse pivot block, upper block U
    using ST_t = ST;
    double T, iT, t5;       // duration, 1/T, T^(1-2S)
    double c[D][3];         // monomial coefficients, ascending powers (row k = c_k)
    ST st;                  // factorisation of this lane's block row
    static __device__ __forceinline__ constexpr int im(int a, int k) { return a * b + k; }
    static __device__ __forceinline__ constexpr int idi(int a, int k) { return (b + a) * b + k; }
    static __device__ __forceinline__ constexpr int iu(int a, int k) { return (2 * b + a) * b + k; }
};
template <int S, int LPT>
using SplineReg = Spline<S, LPT, RegStore<3 * (S - 1) * (S - 1)>>;

// Solve the factorised block-tridiagonal system for a new right-hand side r (in place).
// Row j is  L_j y_{j-1} + D_j y_j + U_j y_{j+1} = r_j.  Forward: r'_j = r_j - M_j r'_{j-1}; backward:
// y_j = Dinv_j (r'_j - U_j y_{j+1}).  Every lane recomputes its row from its own r and its neighbour's
// current value each round, so after round t the rows up to t (from the far end: down to N-1-t) are
// final and later rounds reproduce them: no lane-dependent predicate, `rounds` is warp-uniform.
template <int S, int LPT, class ST>
__device__ __forceinline__ void sweep_apply(unsigned mask, int rounds, const double (&fac)[Spline<S, LPT, ST>::NM],
                                            double (&r)[S - 1][3]) {
    constexpr int b = S - 1;
    using SP = Spline<S, LPT, ST>;
    double rr[b][3];
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) rr[a][x] = r[a][x];
#pragma unroll 1
    for (int t = 0; t < rounds; ++t) {
        double rp[b][3];
#pragma unroll
        for (int a = 0; a < b; ++a)
#pragma unroll
            for (int x = 0; x < 3; ++x) rp[a][x] = sh_up<LPT>(mask, rr[a][x], 1);
#pragma unroll
        for (int a = 0; a < b; ++a)
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                double acc = r[a][x];
#pragma unroll
                for (int k = 0; k < b; ++k) acc = fma(-fac[SP::im(a, k)], rp[k][x], acc);
                rr[a][x] = acc;
            }
    }
    double y[b][3];
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < b; ++k) acc = fma(fac[SP::idi(a, k)], rr[k][x], acc);
            y[a][x] = acc;
        }
#pragma unroll 1
    for (int t = 0; t < rounds; ++t) {
        double yn[b][3], w[b][3];
#pragma unroll
        for (int a = 0; a < b; ++a)
#pragma unroll
            for (int x = 0; x < 3; ++x) yn[a][x] = sh_dn<LPT>(mask, y[a][x], 1);
#pragma unroll
        for (int a = 0; a < b; ++a)
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                double acc = rr[a][x];
#pragma unroll
                for (int k = 0; k < b; ++k) acc = fma(-fac[SP::iu(a, k)], yn[k][x], acc);
                w[a][x] = acc;
            }
#pragma unroll
        for (int a = 0; a < b; ++a)
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < b; ++k) acc = fma(fac[SP::idi(a, k)], w[k][x], acc);
                y[a][x] = acc;
            }
    }
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) r[a][x] = y[a][x];
}

// setParameters: lane `lig` (< N active) holds piece lig with duration T, start position P0,
// end position P1; hd/td are the head / tail derivatives 1..S-1 (used by lanes 0 / N-1).
// Output: sp (coefficients + factorisation), chat (normalised coefficients c_k T^k).
// `rounds` = (pieces of the launch) - 2, the same for every group of the warp (a group that is idle
// passes N = 0 and carries identity rows through the same rounds).
// FRZ (fixed-time mode, MINCOB_FLAG_FREEZE_TIMES): the block factorisation depends on the durations only, which are
// data in that mode.  `refac` (warp-uniform) = factorise and store the multipliers (the first evaluation of a problem);
// otherwise the stored multipliers are loaded and only the right-hand sides are swept (sweep_apply), which performs the
// same operations on the same values: the two paths give identical bits.
template <int S, int LPT, class ST, bool FRZ = false>
__device__ __forceinline__ void spline_solve(unsigned mask, int lig, int N, int rounds, double Tin, const double (&P0)[3],
                                             const double (&P1)[3], const double (&hd)[S - 1][3],
                                             const double (&td)[S - 1][3], Spline<S, LPT, ST> &sp,
                                             double (&chat)[2 * S][3], bool refac = true) {
    constexpr int D = 2 * S, b = S - 1;
    using SP = Spline<S, LPT, ST>;
    using HK = HermiteK<S>;
    const bool active = lig < N;
    const double T = active ? Tin : 1.0;
    const double iT = 1.0 / T;
    double lam[S];
    lam[0] = 1.0;
#pragma unroll
    for (int d = 1; d < S; ++d) lam[d] = lam[d - 1] * T;
    double t5 = iT;
#pragma unroll
    for (int d = 1; d < 2 * S - 1; ++d) t5 *= iT;
    sp.T = T; sp.iT = iT; sp.t5 = t5;
    double dP[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) dP[x] = active ? P1[x] - P0[x] : 0.0;

    // blocks of W(T) = t5 * lam_a lam_c What[a][c]; start unknown rows 1..S-1, end rows S+1..2S-1
    double Bm[b][b], wa[b], wb[b];
#pragma unroll
    for (int a = 0; a < b; ++a) {
        const double la = t5 * lam[a + 1];
#pragma unroll
        for (int c = 0; c < b; ++c) Bm[a][c] = la * lam[c + 1] * HK::W(1 + a, S + 1 + c);
        wa[a] = la * HK::W(1 + a, S);
        wb[a] = la * HK::W(S + 1 + a, S);
    }
    // this piece's contribution to the right-hand side of its end junction (e) and start junction (f)
    double e[b][3], f[b][3];
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            // hd is zero on every lane but 0 and td on every lane but N-1 (all callers pass them so): no predicate needed
            double ev = wb[a] * dP[x], fv = wa[a] * dP[x];
#pragma unroll
            for (int c = 0; c < b; ++c) {
                ev = fma(Bm[c][a], hd[c][x], ev);
                fv = fma(Bm[a][c], td[c][x], fv);
            }
            e[a][x] = ev; f[a][x] = fv;
        }
    // right-hand side of junction `lig` (between piece lig-1 and piece lig)
    double r[b][3];
    const bool junction = (lig >= 1) && (lig < N);
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            const double ep = sh_up<LPT>(mask, e[a][x], 1);
            r[a][x] = junction ? -(ep + f[a][x]) : 0.0;
        }
    double y[b][3];
    if (!FRZ || refac) {
        double A[b][b], C[b][b];
#pragma unroll
        for (int a = 0; a < b; ++a) {
            const double la = t5 * lam[a + 1];
#pragma unroll
            for (int c = 0; c < b; ++c) {
                A[a][c] = la * lam[c + 1] * HK::W(1 + a, 1 + c);
                C[a][c] = la * lam[c + 1] * HK::W(S + 1 + a, S + 1 + c);
            }
        }
        // block row of junction `lig`
        double Dm[b][b], U[b][b], L[b][b];
#pragma unroll
        for (int a = 0; a < b; ++a) {
#pragma unroll
            for (int c = 0; c < b; ++c) {
                const double Cp = sh_up<LPT>(mask, C[a][c], 1);
                const double Bp = sh_up<LPT>(mask, Bm[c][a], 1);  // transposed: L = B_{j-1}^T
                Dm[a][c] = junction ? Cp + A[a][c] : (a == c ? 1.0 : 0.0);
                U[a][c] = (junction && lig < N - 1) ? Bm[a][c] : 0.0;
                L[a][c] = (junction && lig >= 2) ? Bp : 0.0;
            }
        }
        // forward elimination.  M_j = L_j Dinv'_{j-1};  D'_j = D_j - M_j U_{j-1} with U_{j-1} = L_j^T (symmetry);
        // r'_j = r_j - M_j r'_{j-1}.  Each round every lane redoes its row from its ORIGINAL D, r and its
        // neighbour's current values: row j is final after round j-1 and is reproduced unchanged afterwards.
        double Dinv[b][b], M[b][b], rr[b][3];
        inv_small<b>(Dm, Dinv);
#pragma unroll
        for (int a = 0; a < b; ++a) {
#pragma unroll
            for (int c = 0; c < b; ++c) M[a][c] = 0.0;
#pragma unroll
            for (int x = 0; x < 3; ++x) rr[a][x] = r[a][x];
        }
#pragma unroll 1
        for (int t = 0; t < rounds; ++t) {
            double Dp[b][b], rp[b][3];
#pragma unroll
            for (int a = 0; a < b; ++a) {
#pragma unroll
                for (int c = 0; c < b; ++c) Dp[a][c] = sh_up<LPT>(mask, Dinv[a][c], 1);
#pragma unroll
                for (int x = 0; x < 3; ++x) rp[a][x] = sh_up<LPT>(mask, rr[a][x], 1);
            }
            double Dn[b][b];
#pragma unroll
            for (int a = 0; a < b; ++a) {
#pragma unroll
                for (int c = 0; c < b; ++c) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < b; ++k) acc = fma(L[a][k], Dp[k][c], acc);
                    M[a][c] = acc;
                }
            }
#pragma unroll
            for (int a = 0; a < b; ++a) {
#pragma unroll
                for (int c = 0; c < b; ++c) {
                    double acc = Dm[a][c];
#pragma unroll
                    for (int k = 0; k < b; ++k) acc = fma(-M[a][k], L[c][k], acc);
                    Dn[a][c] = acc;
                }
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    double acc = r[a][x];
#pragma unroll
                    for (int k = 0; k < b; ++k) acc = fma(-M[a][k], rp[k][x], acc);
                    rr[a][x] = acc;
                }
            }
            inv_small<b>(Dn, Dinv);
        }
        // back substitution  y_j = Dinv'_j (r'_j - U_j y_{j+1}), same scheme from the far end
#pragma unroll
        for (int a = 0; a < b; ++a) {
#pragma unroll
            for (int k = 0; k < b; ++k) {
                sp.st.put(SP::im(a, k), M[a][k]);
                sp.st.put(SP::idi(a, k), Dinv[a][k]);
                sp.st.put(SP::iu(a, k), U[a][k]);
            }
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < b; ++k) acc = fma(Dinv[a][k], rr[k][x], acc);
                y[a][x] = acc;
            }
        }
#pragma unroll 1
        for (int t = 0; t < rounds; ++t) {
            double yn[b][3], w[b][3];
#pragma unroll
            for (int a = 0; a < b; ++a)
#pragma unroll
                for (int x = 0; x < 3; ++x) yn[a][x] = sh_dn<LPT>(mask, y[a][x], 1);
#pragma unroll
            for (int a = 0; a < b; ++a)
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    double acc = rr[a][x];
#pragma unroll
                    for (int k = 0; k < b; ++k) acc = fma(-U[a][k], yn[k][x], acc);
                    w[a][x] = acc;
                }
#pragma unroll
            for (int a = 0; a < b; ++a)
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < b; ++k) acc = fma(Dinv[a][k], w[k][x], acc);
                    y[a][x] = acc;
                }
        }
    } else {
        // durations unchanged since the multipliers were stored: sweep the right-hand side only
        double fac[SP::NM];
#pragma unroll
        for (int i = 0; i < SP::NM; ++i) fac[i] = sp.st.get(i);
#pragma unroll
        for (int a = 0; a < b; ++a)
#pragma unroll
            for (int x = 0; x < 3; ++x) y[a][x] = r[a][x];
        sweep_apply<S, LPT, ST>(mask, rounds, fac, y);
    }
    // boundary states of this piece, scaled: sh[d] = T^d * (d-th derivative); rows 0..S-1 start,
    // S..2S-1 end (row S holds dP = P1 - P0).  Not kept: spline_adjoint recomputes them from c.
    double sh[D][3];
#pragma unroll
    for (int x = 0; x < 3; ++x) { sh[0][x] = active ? P0[x] : 0.0; sh[S][x] = dP[x]; }
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            const double yn = sh_dn<LPT>(mask, y[a][x], 1);
            const double ys = (lig == 0) ? hd[a][x] : y[a][x];
            const double ye = (lig == N - 1) ? td[a][x] : yn;
            sh[1 + a][x] = lam[a + 1] * ys;       // lanes >= N: identity rows with zero right-hand sides gave y = 0
            sh[S + 1 + a][x] = lam[a + 1] * ye;
        }
    // Hermite -> monomial.  Hhat[k][0] == -Hhat[k][S] for k >= S: positions enter only through dP.
    double ip = 1.0;  // iT^k
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double v;
            if (k < S) {
                v = sh[k][x] * (1.0 / cfact(k));
            } else {
                v = HK::H(k, S) * dP[x];
#pragma unroll
                for (int d = 1; d < S; ++d) v = fma(HK::H(k, S + d), sh[S + d][x], fma(HK::H(k, d), sh[d][x], v));
            }
            chat[k][x] = v;
            sp.c[k][x] = v * ip;
        }
        ip *= iT;
    }
}

// getEnergy + getEnergyPartialGradByCoeffs + getEnergyPartialGradByTimes for this piece
// (SURVEY.md Appendix A.3 in normalised form: E_i = T^(1-2S) chat^T Qhat chat).
// FRZ: the partial derivative by the duration is not wanted (fixed-time mode): gT = 0.
template <int S, int LPT, class ST, bool FRZ = false>
__device__ __forceinline__ void energy_partials(const Spline<S, LPT, ST> &sp, const double (&chat)[2 * S][3], bool active,
                                                double &energy, double (&G)[2 * S][3], double &gT) {
    constexpr int D = 2 * S;
    using HK = HermiteK<S>;
    double e = 0.0, et = 0.0;
    double tp[D];  // T^k
    tp[0] = 1.0;
#pragma unroll
    for (int k = 1; k < D; ++k) tp[k] = tp[k - 1] * sp.T;
#pragma unroll
    for (int k = 0; k < S; ++k)
#pragma unroll
        for (int x = 0; x < 3; ++x) G[k][x] = 0.0;
#pragma unroll
    for (int a = S; a < D; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double qc = 0.0, qt = 0.0;
#pragma unroll
            for (int c = S; c < D; ++c) {
                qc += HK::Q(a, c) * chat[c][x];
                if (!FRZ) qt += (HK::Q(a, c) * (a + c - 2 * S + 1)) * chat[c][x];
            }
            e += chat[a][x] * qc;
            if (!FRZ) et += chat[a][x] * qt;
            G[a][x] = 2.0 * sp.t5 * tp[a] * qc;   // lanes >= N hold zero coefficients (spline_solve): every term is 0 there
        }
    energy = sp.t5 * e;
    gT = FRZ ? 0.0 : sp.t5 * sp.iT * et;
}

// One half-plane row (nx,ny,nz,d); rows are 32-byte aligned.  PSMEM = true: the row sits in the
// group's shared-memory stage (two LDS.128); false: 256-bit read-only global load.
struct __align__(32) Plane { double x, y, z, w; };
struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };
template <bool PSMEM>
__device__ __forceinline__ Plane load_plane(const double *p) {
    Plane r;
    if (PSMEM) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(p);
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(sa));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(r.z), "=d"(r.w) : "r"(sa));
    } else {
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    }
    return r;
}

// Kinematics of one penalty sample: powers of the local time s and the first three derivatives.  The chains start from
// their first term (c_1, c_2, 6 c_3) instead of from zero, and the acceleration is carried as HALF its value
// (hacc = c_2 + 3 s c_3 + 6 s^2 c_4 + ...): scaling by two commutes with every rounding, so 2 * hacc has the bits of the
// directly accumulated sum, and the sample loop saves the doubling (sample_body undoes it for the rare active sample).
template <int S>
__device__ __forceinline__ void sample_kinematics(const double (&c)[2 * S][3], double s, double (&pw)[2 * S], double (&vel)[3],
                                                  double (&hacc)[3], double (&jer)[3]) {
    constexpr int D = 2 * S;
    // derivative bases: bd[k] = k!/(k-d)! s^(k-d)
    pw[0] = 1.0;
#pragma unroll
    for (int k = 1; k < D; ++k) pw[k] = (pw[k - 1] * s);
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        double v = c[1][x], a = c[2][x], jr = 6.0 * c[3][x];
#pragma unroll
        for (int k = 2; k < D; ++k) v = fma((cfall(k, 1) * pw[k - 1]), c[k][x], v);
#pragma unroll
        for (int k = 3; k < D; ++k) a = fma(((0.5 * cfall(k, 2)) * pw[k - 2]), c[k][x], a);
#pragma unroll
        for (int k = 4; k < D; ++k) jr = fma((cfall(k, 3) * pw[k - 3]), c[k][x], jr);
        vel[x] = v; hacc[x] = a; jer[x] = jr;
    }
}
// |v|^2 - v_max^2 and friends (> 0: the hinge is active)
__device__ __forceinline__ double excess(const double (&u)[3], double lim2) {
    return fma(u[2], u[2], fma(u[1], u[1], fma(u[0], u[0], -lim2)));
}

// Contribution of ONE active penalty sample j to cost, dF/dc (G) and the partial dF/dT (gT): the hinges, their
// gradients and the chain rule through the sample's bases.  hp: some half-plane test of this sample was not strictly
// negative in phase 1; pm0/pm1: rows to re-test exactly (any superset of the rows that can be positive).
// This is the only place where the penalty accumulates, one sample after the other in ascending j -- in both mappings.
template <int S, int PSM, bool FRZ>
__device__ __forceinline__ void sample_body(const DevParams &P, const double (&c)[2 * S][3], const double *planes, int rstride,
                                            unsigned pm0, unsigned pm1, bool hp, int j, int kap, double ikap, double imu,
                                            double step, const double (&pw)[2 * S], const double (&vel)[3],
                                            const double (&hacc)[3], const double (&jer)[3], double vv, double aa4, double jj2,
                                            double &cost, double (&G)[2 * S][3], double &gT) {
    constexpr int D = 2 * S;
    // sample_kinematics carries half the acceleration and the caller a quarter of its excess: exact to undo
    const double acc[3] = {2.0 * hacc[0], 2.0 * hacc[1], 2.0 * hacc[2]};
    const double aa = 4.0 * aa4;
    double pena = 0.0, gP[3] = {0, 0, 0}, gV[3] = {0, 0, 0}, gA[3] = {0, 0, 0}, gJ[3] = {0, 0, 0};
    double fv, df;
    if (hp) {
        double pos[3];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double v = 0.0;
#pragma unroll
            for (int k = D - 1; k >= 0; --k) v = fma(pw[k], c[k][x], v);
            pos[x] = v;
        }
        // only rows flagged for this block (in row order, as the reference sums them)
#pragma unroll 1
        for (int w = 0; w < 2; ++w) {
            unsigned todo = w ? pm1 : pm0;
#pragma unroll 1
            while (todo != 0u) {
                const int k = 32 * w + __ffs(todo) - 1;
                todo &= todo - 1u;
                const Plane h = load_plane<PSM == 1>(planes + (size_t)k * rstride);
                const double v = fma(h.x, pos[0], fma(h.y, pos[1], fma(h.z, pos[2], h.w)));
                if (v > 0.0) {
                    smoothed_l1_pos(P.mu, imu, v, fv, df);
                    const double wd = (P.w_pos * df);
                    gP[0] = fma(wd, h.x, gP[0]); gP[1] = fma(wd, h.y, gP[1]); gP[2] = fma(wd, h.z, gP[2]);
                    pena = fma(P.w_pos, fv, pena);
                }
            }
        }
    }
    if (vv > 0.0) {
        smoothed_l1_pos(P.mu, imu, vv, fv, df);
        const double wd = ((P.w_vel * df) * 2.0);
        gV[0] = (wd * vel[0]); gV[1] = (wd * vel[1]); gV[2] = (wd * vel[2]);
        pena = fma(P.w_vel, fv, pena);
    }
    if (aa > 0.0) {
        smoothed_l1_pos(P.mu, imu, aa, fv, df);
        const double wd = ((P.w_acc * df) * 2.0);
        gA[0] = (wd * acc[0]); gA[1] = (wd * acc[1]); gA[2] = (wd * acc[2]);
        pena = fma(P.w_acc, fv, pena);
    }
    if (jj2 > 0.0) {
        smoothed_l1_pos(P.mu, imu, jj2, fv, df);
        const double wd = ((P.w_jerk * df) * 2.0);
        gJ[0] = (wd * jer[0]); gJ[1] = (wd * jer[1]); gJ[2] = (wd * jer[2]);
        pena = fma(P.w_jerk, fv, pena);
    }
    const double node = (j == 0 || j == kap) ? 0.5 : 1.0;
    const double w = (node * step);
    double dsum = 0.0;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        if (!FRZ) {
            double sn = 0.0;
#pragma unroll
            for (int k = 4; k < D; ++k) sn = fma((cfall(k, 4) * pw[k - 4]), c[k][x], sn);
            dsum = fma(gJ[x], sn, fma(gA[x], jer[x], fma(gV[x], acc[x], fma(gP[x], vel[x], dsum))));
        }
        const double wP = (w * gP[x]), wV = (w * gV[x]), wA = (w * gA[x]), wJ = (w * gJ[x]);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double t = (pw[k] * wP);
            if (k >= 1) t = fma((cfall(k, 1) * pw[k - 1]), wV, t);
            if (k >= 2) t = fma((cfall(k, 2) * pw[k - 2]), wA, t);
            if (k >= 3) t = fma((cfall(k, 3) * pw[k - 3]), wJ, t);
            G[k][x] = __dadd_rn(G[k][x], t);
        }
    }
    if (!FRZ) {
        const double through_t = ((dsum * ((double)j * ikap)) * w);
        gT = __dadd_rn(gT, fma((node * pena), ikap, through_t));
    }
    cost = fma(w, pena, cost);
}

// attachPenaltyFunctional for this piece.  planes: this piece's rows.
// Phase 1 tests a register block of JB sample positions against every half-plane, so each row is
// loaded once per block instead of once per sample.  Phase 2 is a ROLLED loop over the samples of
// the block (the position is recomputed, 15 DFMA, rather than indexed out of registers): the hot
// loop has to stay inside the 32 KB instruction cache.
// Rows touched by a block are remembered in a 64-bit mask (two words: the second one is only ever written when
// K > 32, e.g. the reference's polytopes padded to 50 rows, learning_planner.hpp:40,157-168); K <= MINCOB_MAX_ROWS = 64.
// REP (latency mapping, all lanes of the warp call this, `live` = the lane holds a piece): the replicas split the TESTS --
// this lane runs phase 1 and the hinge tests of the samples j = jbase + u * jstride only and records which of them are
// active; the flag words are ORed over the replicas, and then EVERY replica accumulates the active samples itself, in
// ascending j like the throughput mapping (sample_body; the replicas hold the same coefficients, so nothing but flags
// is exchanged).  The sums are therefore bit-identical to REP = false, and the replicas stay bit-identical to each other.
// Needs kappa <= 31 (one flag bit per sample; the launcher checks).
// PSM: where the rows are.  0: global memory (C-ABI layout); 1: the group's shared-memory stage.  (A hybrid -- the
// first rows of polytopes too large to stage whole in the stage, the rest in global memory -- was built and measured in
// round 2: no gain at any split, profiles/r02_tail_experiments.md section 7.)
// FRZ (fixed-time mode): the time-gradient terms (through the sample times and the quadrature weights) are not computed.
template <int S, int LPT, int PSM, bool REP, class ST, bool FRZ = false>
__device__ __forceinline__ void penalty_piece(const DevParams &P, const Spline<S, LPT, ST> &sp, const double *planes,
                                              int rstride, int K, bool live, int jbase, int jstride, double &cost,
                                              double (&G)[2 * S][3], double &gT) {
    constexpr int D = 2 * S, JB = MINCOB_JB;
    constexpr bool DEFER = REP || MINCOB_DEFER_BODY;   // active samples are accumulated after all tests, not inside the test loop
    const int kap = P.kappa;
    const double ikap = P.ikap, imu = P.imu;
    const double step = sp.T * ikap;
    const int cnt = REP ? ((live && jbase <= kap) ? (kap - jbase) / jstride + 1 : 0)
                        : ((MINCOB_DEFER_BODY && !live) ? 0 : kap + 1);   // samples this lane tests
    auto sample = [&](int u) { return REP ? jbase + u * jstride : u; };
    unsigned need = 0u, hitj = 0u, rm0 = 0u, rm1 = 0u;   // REP: active samples / samples with a flagged row (bit j), flagged rows
#pragma unroll 1
    for (int j0 = 0; j0 < cnt; j0 += JB) {
        unsigned hit = 0u, pm0 = 0u, pm1 = 0u;
        {
            // sign-bit test: keep[jj] stays negative only while every n.p + d is negative; a sample
            // with any value >= +0 is flagged and re-tested exactly (v > 0) in phase 2.  (Dropping the per-sample masks and
            // working the flagged samples out of the flagged rows afterwards executes fewer instructions and is 2-3 %
            // slower: profiles/r02_ab.log.)
            int keep[JB];
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) keep[jj] = -1;
            double pos[JB][3];
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) {
                // (slots past this lane's last sample repeat it: a position beyond the end of the piece would flag rows)
                const double s = sample(min(j0 + jj, cnt - 1)) * step;
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    double v = sp.c[D - 1][x];
#pragma unroll
                    for (int k = D - 2; k >= 0; --k) v = fma(v, s, sp.c[k][x]);
                    pos[jj][x] = v;
                }
            }
            // rows in shared memory: one per trip (code size); rows in global memory (K too large to stage): several
            // loads in flight per trip, the L2 latency is what that loop waits for
            auto scan = [&](auto in_smem, const double *base, int stride, int ka, int kb, unsigned &pm) {
                constexpr bool SM = decltype(in_smem)::value;
                constexpr int UK = SM ? MINCOB_UNROLL_K : MINCOB_UNROLL_KG;
                auto test_row = [&](const Plane &h, unsigned bit) {
                    int all = -1;
#pragma unroll
                    for (int jj = 0; jj < JB; ++jj) {
                        const double v = fma(h.x, pos[jj][0], fma(h.y, pos[jj][1], fma(h.z, pos[jj][2], h.w)));
                        keep[jj] &= __double2hiint(v);
                        all &= __double2hiint(v);
                    }
                    pm |= all < 0 ? 0u : bit;   // row touched by some sample of the block
                };
#if MINCOB_PLANE_PREFETCH
                if constexpr (SM) {
                    // software pipelining by one row: the next row's load is in flight while this row's tests issue.  In
                    // the stage the row after a polytope's last one is still inside the group's shared memory (the next
                    // rows or the small arrays behind the stage), so the pointer just advances: no clamp, no multiply.
                    const double *p = base + (size_t)ka * stride;
                    Plane hnext = load_plane<true>(p);
                    unsigned bit = 1u << (ka & 31);
#pragma unroll UK
                    for (int k = ka; k < kb; ++k) {
                        const Plane h = hnext;
                        p += stride;
                        hnext = load_plane<true>(p);
                        test_row(h, bit);
                        bit <<= 1;
                    }
                } else {
                    Plane hnext = load_plane<false>(base + (size_t)(ka < kb ? ka : 0) * stride);
#pragma unroll UK
                    for (int k = ka; k < kb; ++k) {
                        const Plane h = hnext;
                        hnext = load_plane<false>(base + (size_t)(k + 1 < kb ? k + 1 : k) * stride);
                        test_row(h, 1u << (k & 31));
                    }
                }
#else
#pragma unroll UK
                for (int k = ka; k < kb; ++k) test_row(load_plane<SM>(base + (size_t)k * stride), 1u << (k & 31));
#endif
            };
#pragma unroll 1
            for (int k0 = 0; k0 < K; k0 += 32) {      // one mask word per trip; a single trip unless K > 32
                unsigned pm = 0u;
                const int k1 = min(K, k0 + 32);
                if constexpr (PSM == 1) scan(TrueT(), planes, rstride, k0, k1, pm);
                else scan(FalseT(), planes, rstride, k0, k1, pm);
                if (k0 == 0) pm0 = pm; else pm1 = pm;
            }
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) hit |= (keep[jj] < 0 ? 0u : 1u) << jj;
        }
        const int jend = min(JB, cnt - j0);
        constexpr int UJ = MINCOB_UNROLL_JJ;

#pragma unroll UJ
        for (int jj = 0; jj < jend; ++jj) {
            const int j = sample(j0 + jj);
            const double s = j * step;
            double pw[D], vel[3], acc[3], jer[3];
            sample_kinematics<S>(sp.c, s, pw, vel, acc, jer);   // acc: half the acceleration
            // aa: a quarter of |a|^2 - a_max^2 (same sign, exactly a quarter of the value)
            const double vv = excess(vel, P.vmax2), aa = excess(acc, P.amax2q), jj2 = excess(jer, P.jmax2);
            const bool hp = (hit >> jj) & 1u;
            if (hp || vv > 0.0 || aa > 0.0 || jj2 > 0.0) {
                if constexpr (DEFER) {
                    need |= 1u << j;
                    hitj |= (hp ? 1u : 0u) << j;
                } else {
                    sample_body<S, PSM, FRZ>(P, sp.c, planes, rstride, pm0, pm1, hp, j, kap, ikap, imu, step, pw, vel, acc, jer,
                                             vv, aa, jj2, cost, G, gT);
                }
            }
        }
        if constexpr (DEFER) { rm0 |= pm0; rm1 |= pm1; }
    }
    if constexpr (DEFER) {
        constexpr unsigned FULL = 0xffffffffu;
        if constexpr (REP) {
            need = replica_or<LPT>(FULL, need);
            hitj = replica_or<LPT>(FULL, hitj);
            rm0 = replica_or<LPT>(FULL, rm0);
            rm1 = replica_or<LPT>(FULL, rm1);
        }
        if (!live) need = 0u;
        // every lane works through ITS active samples in ascending order: the trips are the largest count of any lane
#pragma unroll 1
        while (__any_sync(FULL, need != 0u)) {
            if (need != 0u) {
                const int j = __ffs(need) - 1;
                need &= need - 1u;
                const double s = j * step;
                double pw[D], vel[3], acc[3], jer[3];
                sample_kinematics<S>(sp.c, s, pw, vel, acc, jer);
                const double vv = excess(vel, P.vmax2), aa = excess(acc, P.amax2q), jj2 = excess(jer, P.jmax2);
                sample_body<S, PSM, FRZ>(P, sp.c, planes, rstride, rm0, rm1, (hitj >> j) & 1u, j, kap, ikap, imu, step, pw, vel,
                                         acc, jer, vv, aa, jj2, cost, G, gT);
            }
        }
    }
}

// propogateGrad: G = dF/dc_i, gTp = partial dF/dT_i  ->  total dJ/dq_lig (junction lig, lanes
// 1..N-1) and dJ/dT_lig (lanes 0..N-1).  See oracle/reduced_proto.py for the derivation.
// FRZ (fixed-time mode): only dJ/dq is computed (gT = 0): the boundary states, What (L s) and the m^T W' s terms drop out.
template <int S, int LPT, class ST, bool FRZ = false>
__device__ __forceinline__ void spline_adjoint(unsigned mask, int lig, int N, int rounds, const Spline<S, LPT, ST> &sp,
                                               const double (&G)[2 * S][3], double gTp, double (&gq)[3], double &gT) {
    constexpr int D = 2 * S, b = S - 1;
    using HK = HermiteK<S>;
    const bool active = lig < N;
    const bool junction = (lig >= 1) && active;
    // the factorisation of this lane's block row, requested first: with GlobalStore these are independent L2 loads
    // whose latency the change of basis below covers
    double mul[Spline<S, LPT, ST>::NM];
#pragma unroll
    for (int i = 0; i < Spline<S, LPT, ST>::NM; ++i) mul[i] = sp.st.get(i);
    // ghat_k = G_k / T^k ; z = Hhat^T ghat ; through-H time term  -(1/T) sum_k k G_k.c_k
    double gh[D][3];
    double ip = 1.0, kGc = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            gh[k][x] = G[k][x] * ip;
            if (!FRZ) kGc += (double)k * G[k][x] * sp.c[k][x];
        }
        ip *= sp.iT;
    }
    double z[D][3];
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double v = (d < S) ? gh[d][x] * (1.0 / cfact(d)) : 0.0;
#pragma unroll
            for (int k = S; k < D; ++k) v += HK::H(k, d) * gh[k][x];
            z[d][x] = v;
        }
    double lam[S];
    lam[0] = 1.0;
#pragma unroll
    for (int d = 1; d < S; ++d) lam[d] = lam[d - 1] * sp.T;
    // scaled boundary states of this piece from its coefficients (chat_k = c_k T^k):
    //   start sh[d] = d! chat_d,  end sh[S+d] = sum_k k!/(k-d)! chat_k,  sh[S] = dP = sum_{k>=1} chat_k
    double sh[D][3];
    if (!FRZ) {
        double tk = 1.0;
        double chat[D][3];
#pragma unroll
        for (int k = 0; k < D; ++k) {
#pragma unroll
            for (int x = 0; x < 3; ++x) chat[k][x] = sp.c[k][x] * tk;   // zero on lanes >= N, like G and hence z
            tk *= sp.T;
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) {
#pragma unroll
            for (int d = 0; d < S; ++d) {
                sh[d][x] = cfact(d) * chat[d][x];
                double e = 0.0;
#pragma unroll
                for (int k = D - 1; k >= (d == 0 ? 1 : d); --k) e += cfall(k, d) * chat[k][x];
                sh[S + d][x] = e;
            }
        }
    }
    // gather at junctions: g_y[j] = (L z)_{end, piece j-1} + (L z)_{start, piece j}
    double r[b][3], gp[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        const double pe = sh_up<LPT>(mask, z[S][x], 1);
        gp[x] = junction ? pe + z[0][x] : 0.0;
    }
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            const double ee = sh_up<LPT>(mask, lam[a + 1] * z[S + 1 + a][x], 1);
            r[a][x] = junction ? ee + lam[a + 1] * z[1 + a][x] : 0.0;
        }
    sweep_apply<S, LPT, ST>(mask, rounds, mul, r);  // r <- mu_lig
    // m_i = [0, mu_i ; 0, mu_{i+1}], scaled by L
    double lm[D][3];
#pragma unroll
    for (int x = 0; x < 3; ++x) { lm[0][x] = 0.0; lm[S][x] = 0.0; }
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            const double mn = sh_dn<LPT>(mask, r[a][x], 1);
            lm[1 + a][x] = lam[a + 1] * r[a][x];   // r = 0 off the junctions (identity rows)
            lm[S + 1 + a][x] = (lig < N - 1) ? lam[a + 1] * mn : 0.0;
        }
    // wm = What (L m), ws = What (L s)  (positions through dP: What[:,0] == -What[:,S])
    double acc_ms = 0.0, acc_dms = 0.0, acc_dsm = 0.0, zds = 0.0;
    double wm0[3], wmS[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        double wm[D], ws[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            if (FRZ && a != 0 && a != S) continue;   // only rows 0 and S of What (L m) reach dJ/dq
            double vm = 0.0, vs = FRZ ? 0.0 : HK::W(a, S) * sh[S][x];
#pragma unroll
            for (int d = 1; d < S; ++d) {
                vm = fma(HK::W(a, S + d), lm[S + d][x], fma(HK::W(a, d), lm[d][x], vm));
                if (!FRZ) vs = fma(HK::W(a, S + d), sh[S + d][x], fma(HK::W(a, d), sh[d][x], vs));
            }
            wm[a] = vm; ws[a] = vs;
        }
        wm0[x] = wm[0]; wmS[x] = wm[S];
        if (!FRZ) {
#pragma unroll
            for (int d = 1; d < S; ++d) {
                const double lw = fma(lm[S + d][x], ws[S + d], lm[d][x] * ws[d]);
                acc_ms += lw;
                acc_dms = fma((double)d, lw, acc_dms);
                acc_dsm = fma((double)d, fma(sh[S + d][x], wm[S + d], sh[d][x] * wm[d]), acc_dsm);
                zds = fma((double)d, fma(z[S + d][x], sh[S + d][x], z[d][x] * sh[d][x]), zds);
            }
        }
    }
    // dJ/dq_j = g_p[j] - t5_j wm_j[0] - (t5 wm[S])_{j-1}
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        const double prev = sh_up<LPT>(mask, sp.t5 * wmS[x], 1);
        gq[x] = junction ? gp[x] - sp.t5 * wm0[x] - prev : 0.0;
    }
    // dJ/dT_i = partial - (1/T) sum k G_k.c_k + z.(L' s) - m^T W' s ;  L' = (d/T) L
    const double mWs = sp.t5 * sp.iT * (-(double)(2 * S - 1) * acc_ms + acc_dms + acc_dsm);
    gT = (active && !FRZ) ? gTp + sp.iT * (zds - kGc) - mWs : 0.0;
}

// ---- problem view ---------------------------------------------------------------------
// What one group needs of its problem.  `planes` points at THIS LANE's first row and `rstride` is the
// distance (in doubles) between consecutive rows of the lane's polytope:
//   C-ABI layout in global memory  [N][K][4]:  planes = base + lig*K*4, rstride = 4
//   shared-memory stage, plane-major [K][N][4]: planes = base + lig*4,   rstride = 4*N
//     (for one k the lanes of a group read 32*N contiguous bytes: conflict-free LDS.128)
struct ProblemView {
    const double *head;    // [S][3]
    const double *tail;    // [S][3]
    const double *planes;  // or nullptr
    int rstride;
    int rows;              // rows of this lane's polytope actually used (<= K)
};

// Whole cost functional for the group's trajectory.  xt = tau_lig, xq = q_lig (lanes 1..N-1).
// Returns f on every lane of the group; gt = dJ/dtau_lig, gq = dJ/dq_lig.
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// `before_adjoint` runs between the penalty loop and the adjoint: the optimize kernel uses it to request
// its parked optimizer state early (plain loads whose latency the adjoint then covers).
// REP: the groups of the warp are replicas of one trajectory (latency mapping): each tests every GPW-th penalty sample,
// the flags are exchanged, and every replica accumulates the active samples in ascending order (penalty_piece): the
// result has the bits of the throughput mapping.
// FRZ: fixed-time specialisation (P.freeze set by the caller): the stored block factorisation is reused unless `refac`
// (warp-uniform: some group of the warp evaluates a newly fetched problem), and nothing of dJ/dT is computed.
template <int S, int LPT, int PSM, class ST, bool REP = false, class Hook = NoHook, bool FRZ = false>
__device__ __forceinline__ double cost_functional(const DevParams &P, unsigned mask, int lig, int N, int rounds,
                                                  const ProblemView &pv, const ST &store, double xt,
                                                  const double (&xq)[3], double &gt, double (&gq)[3],
                                                  Hook before_adjoint = Hook(), bool refac = true) {
    constexpr int D = 2 * S, b = S - 1;
    const bool active = lig < N;
    double P0[3], P1[3], hd[b][3], td[b][3];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        const double qn = sh_dn<LPT>(mask, xq[x], 1);
        P0[x] = (lig == 0) ? pv.head[x] : xq[x];
        P1[x] = (lig == N - 1) ? pv.tail[x] : qn;
    }
#pragma unroll
    for (int a = 0; a < b; ++a)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            hd[a][x] = (lig == 0) ? pv.head[(a + 1) * 3 + x] : 0.0;
            td[a][x] = (lig == N - 1) ? pv.tail[(a + 1) * 3 + x] : 0.0;
        }
    const double T = active ? forward_t(xt) : 1.0;
    Spline<S, LPT, ST> sp;
    sp.st = store;
    double chat[D][3];
    constexpr bool FS = FRZ, FT = FRZ;   // reuse of the factorisation / no time-gradient terms
    spline_solve<S, LPT, ST, FS>(mask, lig, N, rounds, T, P0, P1, hd, td, sp, chat, refac);
    double cost, G[D][3], gTp;
    energy_partials<S, LPT, ST, FT>(sp, chat, active, cost, G, gTp);
    // (REP: every replica computes the energy terms itself and accumulates the same active samples in the same order:
    // nothing to add up afterwards)
    if (P.penalties) {
        if constexpr (REP) {
            penalty_piece<S, LPT, PSM, true, ST, FT>(P, sp, pv.planes, pv.rstride, (active && pv.planes) ? pv.rows : 0, active,
                                                     Lanes<LPT>::giw(), Lanes<LPT>::GPW, cost, G, gTp);
        } else if constexpr (MINCOB_DEFER_BODY) {   // (all lanes call: the deferred accumulation loop votes over the warp)
            penalty_piece<S, LPT, PSM, false, ST, FT>(P, sp, pv.planes, pv.rstride, (active && pv.planes) ? pv.rows : 0, active,
                                                      0, 1, cost, G, gTp);
        } else if (active) {
            penalty_piece<S, LPT, PSM, false, ST, FT>(P, sp, pv.planes, pv.rstride, pv.planes ? pv.rows : 0, true, 0, 1,
                                                      cost, G, gTp);
        }
    }
    before_adjoint();
    double gT;
    spline_adjoint<S, LPT, ST, FT>(mask, lig, N, rounds, sp, G, gTp, gq, gT);
    if (active) cost = fma(P.rho, T, cost);
    // freeze: durations are data, not variables.  Deliberately a run-time test also in the FRZ instantiation (where it is
    // always true): with a compile-time zero the compiler folds it into the L-BFGS dot products and contracts them
    // differently, and the fixed-time kernel would no longer reproduce the generic kernel bit for bit (measured).
    gt = (active && !P.freeze) ? backward_grad_t(xt, gT + P.rho) : 0.0;
    return group_sum<LPT>(mask, cost);
}

}  // namespace mincob
