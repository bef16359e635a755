// ============================================================================
// oracle/minco_oracle.hpp — CPU fp64 ORACLE for the MINCO hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under allocnet_b200/ or include/ may
// include, link or execute this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it (as the checker / CPU
// baseline, never as the product path).
//
// What it restates (plain C++17, no Eigen, no third-party code):
//   * BandedSystem / MINCO_S3NU / MINCO_S4NU of upstream ZJU-FAST-Lab/GCOPTER
//     (gcopter/include/gcopter/minco.hpp).  That module is NOT vendored in
//     /root/reference (.SUBMODULES.json "submodules": []), so
//     there is no pinned version; the published algorithm is restated from
//     SURVEY.md Appendix A and anchored on the reference's own call sites:
//       - variable / coefficient layout  planner/qp_solver.hpp:133,166-169
//       - continuity + boundary rows     planner/qp_solver.hpp:148-177
//       - jerk-energy Q constants        planner/qp_solver.hpp:223-234,
//                                        gcopter/trajectory.hpp:396-420
//       - Piece<D> descending powers     gcopter/trajectory.hpp:75-133
//   * GCOPTER_PolytopeSFC::costFunctional / attachPenaltyFunctional /
//     forwardT / backwardT / backwardGradT (upstream gcopter.hpp; SURVEY.md
//     Appendix B) with the smoothed hinge of gcopter/firi.hpp:60-84.
//
// PARITY STATUS: the reference holds no tests, golden vectors or fixtures for
// this path and its MINCO code is absent => the MINCO /
// cost-functional part of this oracle is "parity unpinned" by reference
// artefacts.  It is pinned instead by (tests/test_oracle_*.py): finite
// differences, an independent dense numpy/mpmath restatement, the KKT
// equivalence with the reference's own QP formulation (qp_solver.hpp rows and
// Q), and E == 2*getTrajCost(3) (trajectory.hpp:396-420).  The L-BFGS part
// (oracle/lbfgs_oracle.hpp) IS pinned against the real reference: lbfgs.hpp is
// compiled verbatim from /root/reference into oracle/_ref (see oracle/Makefile).
// ============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------
// Banded linear system, diagonal-major storage, no-pivot LU.
// Follows upstream minco.hpp BandedSystem:
// element (i,j) lives at data[(i - j + q) * n + j].
// ---------------------------------------------------------------------------
class Banded {
public:
    void create(int n_, int p_, int q_) {
        n = n_; p = p_; q = q_;
        data.assign(static_cast<size_t>(n) * (p + q + 1), 0.0);
    }
    void reset() { std::fill(data.begin(), data.end(), 0.0); }
    inline double &at(int i, int j) { return data[static_cast<size_t>(i - j + q) * n + j]; }
    inline double at(int i, int j) const { return data[static_cast<size_t>(i - j + q) * n + j]; }
    int size() const { return n; }
    int lower() const { return p; }
    int upper() const { return q; }

    // In-place LU without pivoting (Golub & Van Loan band variant).
    void factorizeLU() {
        for (int k = 0; k <= n - 2; ++k) {
            const int iM = std::min(k + p, n - 1);
            double piv = at(k, k);
            for (int i = k + 1; i <= iM; ++i)
                if (at(i, k) != 0.0) at(i, k) /= piv;
            const int jM = std::min(k + q, n - 1);
            for (int j = k + 1; j <= jM; ++j) {
                const double ukj = at(k, j);
                if (ukj != 0.0)
                    for (int i = k + 1; i <= iM; ++i)
                        if (at(i, k) != 0.0) at(i, j) -= at(i, k) * ukj;
            }
        }
    }
    // b is n x m, row-major (row stride m); overwritten by A^{-1} b.
    void solve(double *b, int m) const {
        for (int j = 0; j <= n - 1; ++j) {
            const int iM = std::min(j + p, n - 1);
            for (int i = j + 1; i <= iM; ++i) {
                const double l = at(i, j);
                if (l != 0.0)
                    for (int c = 0; c < m; ++c) b[i * m + c] -= l * b[j * m + c];
            }
        }
        for (int j = n - 1; j >= 0; --j) {
            const double d = at(j, j);
            for (int c = 0; c < m; ++c) b[j * m + c] /= d;
            const int iM = std::max(0, j - q);
            for (int i = iM; i <= j - 1; ++i) {
                const double u = at(i, j);
                if (u != 0.0)
                    for (int c = 0; c < m; ++c) b[i * m + c] -= u * b[j * m + c];
            }
        }
    }
    // b overwritten by A^{-T} b.
    void solveAdj(double *b, int m) const {
        for (int j = 0; j <= n - 1; ++j) {
            const double d = at(j, j);
            for (int c = 0; c < m; ++c) b[j * m + c] /= d;
            const int iM = std::min(j + q, n - 1);
            for (int i = j + 1; i <= iM; ++i) {
                const double u = at(j, i);
                if (u != 0.0)
                    for (int c = 0; c < m; ++c) b[i * m + c] -= u * b[j * m + c];
            }
        }
        for (int j = n - 1; j >= 0; --j) {
            const int iM = std::max(0, j - p);
            for (int i = iM; i <= j - 1; ++i) {
                const double l = at(j, i);
                if (l != 0.0)
                    for (int c = 0; c < m; ++c) b[i * m + c] -= l * b[j * m + c];
            }
        }
    }

private:
    int n = 0, p = 0, q = 0;
    std::vector<double> data;
};

// d-th derivative of the monomial basis at t: out[k] = k!/(k-d)! t^(k-d), k<D.
template <int D>
inline void basis(double t, int d, double *out) {
    for (int k = 0; k < D; ++k) {
        if (k < d) { out[k] = 0.0; continue; }
        double f = 1.0;
        for (int u = 0; u < d; ++u) f *= static_cast<double>(k - u);
        double tp = 1.0;
        for (int u = 0; u < k - d; ++u) tp *= t;
        out[k] = f * tp;
    }
}
// rows d = 0..ND-1 of the derivative basis at t from ONE table of powers: pw[k] is built by the same repeated
// multiplication as in basis(), so every entry has the same bits; the factors k!/(k-d)! are exact small integers.
template <int D, int ND>
inline void basis_rows(double t, double (&out)[ND][D]) {
    double pw[D];
    pw[0] = 1.0;
    for (int k = 1; k < D; ++k) pw[k] = pw[k - 1] * t;
    for (int d = 0; d < ND; ++d)
        for (int k = 0; k < D; ++k) {
            if (k < d) { out[d][k] = 0.0; continue; }
            double f = 1.0;
            for (int u = 0; u < d; ++u) f *= static_cast<double>(k - u);
            out[d][k] = f * pw[k - d];
        }
}
inline double factorial(int d) {
    double f = 1.0;
    for (int u = 2; u <= d; ++u) f *= u;
    return f;
}

// ---------------------------------------------------------------------------
// MINCO, order S (S=3: MINCO_S3NU, quintic pieces, jerk energy;
//                 S=4: MINCO_S4NU, septic pieces, snap energy).
// Layouts (all double, the Eigen column-major buffers of the upstream API):
//   head/tail : 3 x S  col-major  -> [d*3 + axis]   (P,V,A[,J])
//   inPs      : 3 x (N-1) col-major -> [i*3 + axis]
//   coeffs b  : (2S*N) x 3, stored ROW-major here: b[(2S*i + k)*3 + axis],
//               ascending powers k.
// ---------------------------------------------------------------------------
template <int S>
class Minco {
public:
    static constexpr int D = 2 * S;

    void setConditions(const double *headState, const double *tailState, int pieceNum) {
        N = pieceNum;
        std::memcpy(head, headState, sizeof(double) * 3 * S);
        std::memcpy(tail, tailState, sizeof(double) * 3 * S);
        A.create(D * N, D, D);
        b.assign(static_cast<size_t>(D) * N * 3, 0.0);
        T.assign(N, 0.0);
    }

    void setParameters(const double *inPs, const double *ts) {
        for (int i = 0; i < N; ++i) T[i] = ts[i];
        A.reset();
        std::fill(b.begin(), b.end(), 0.0);
        double beta[D];
        // head rows: derivative d at t=0.
        for (int d = 0; d < S; ++d) {
            A.at(d, d) = factorial(d);
            for (int a = 0; a < 3; ++a) b[d * 3 + a] = head[d * 3 + a];
        }
        for (int i = 0; i < N - 1; ++i) {
            const int c0 = D * i;            // first column of piece i
            const int r = D * i + S;         // first row of junction i
            // continuity of derivatives S..2S-2
            for (int j = 0; j <= S - 2; ++j) {
                const int d = S + j;
                basis<D>(T[i], d, beta);
                for (int k = d; k < D; ++k) A.at(r + j, c0 + k) = beta[k];
                A.at(r + j, c0 + D + d) = -factorial(d);
            }
            // waypoint row
            basis<D>(T[i], 0, beta);
            for (int k = 0; k < D; ++k) A.at(r + S - 1, c0 + k) = beta[k];
            for (int a = 0; a < 3; ++a) b[(r + S - 1) * 3 + a] = inPs[i * 3 + a];
            // continuity of derivatives 0..S-1
            for (int d = 0; d < S; ++d) {
                basis<D>(T[i], d, beta);
                for (int k = d; k < D; ++k) A.at(r + S + d, c0 + k) = beta[k];
                A.at(r + S + d, c0 + D + d) = -factorial(d);
            }
        }
        // tail rows
        for (int d = 0; d < S; ++d) {
            basis<D>(T[N - 1], d, beta);
            const int row = D * N - S + d;
            for (int k = d; k < D; ++k) A.at(row, D * (N - 1) + k) = beta[k];
            for (int a = 0; a < 3; ++a) b[row * 3 + a] = tail[d * 3 + a];
        }
        A.factorizeLU();
        A.solve(b.data(), 3);
    }

    const std::vector<double> &getCoeffs() const { return b; }
    const std::vector<double> &getTimes() const { return T; }
    int pieces() const { return N; }

    // E = sum_i int_0^Ti |p^(S)|^2 dt  (no 1/2).
    // T^e, e = 0 .. 2S-1, by repeated multiplication (the timed CPU path should not pay for std::pow)
    static void tpowers(double t, double (&tp)[D]) {
        tp[0] = 1.0;
        for (int e = 1; e < D; ++e) tp[e] = tp[e - 1] * t;
    }
    void getEnergy(double &energy) const {
        energy = 0.0;
        double tp[D];
        for (int i = 0; i < N; ++i) {
            tpowers(T[i], tp);
            for (int a = S; a < D; ++a)
                for (int c = S; c < D; ++c) {
                    const int e = a + c - 2 * S + 1;
                    const double m = fallfac(a) * fallfac(c) * tp[e] / e;
                    energy += m * dot3(&b[(D * i + a) * 3], &b[(D * i + c) * 3]);
                }
        }
    }
    // gdC: (2S*N) x 3 row-major, OVERWRITTEN.
    void getEnergyPartialGradByCoeffs(double *gdC) const {
        std::fill(gdC, gdC + static_cast<size_t>(D) * N * 3, 0.0);
        double tp[D];
        for (int i = 0; i < N; ++i) {
            tpowers(T[i], tp);
            for (int a = S; a < D; ++a)
                for (int c = S; c < D; ++c) {
                    const int e = a + c - 2 * S + 1;
                    const double m = 2.0 * fallfac(a) * fallfac(c) * tp[e] / e;
                    for (int x = 0; x < 3; ++x)
                        gdC[(D * i + a) * 3 + x] += m * b[(D * i + c) * 3 + x];
                }
        }
    }
    // gdT: N, OVERWRITTEN.
    void getEnergyPartialGradByTimes(double *gdT) const {
        double tp[D];
        for (int i = 0; i < N; ++i) {
            double g = 0.0;
            tpowers(T[i], tp);
            for (int a = S; a < D; ++a)
                for (int c = S; c < D; ++c) {
                    const int e = a + c - 2 * S + 1;
                    const double m = fallfac(a) * fallfac(c) * tp[e - 1];
                    g += m * dot3(&b[(D * i + a) * 3], &b[(D * i + c) * 3]);
                }
            gdT[i] = g;
        }
    }

    // Adjoint: maps partial dJ/dc, dJ/dT to total dJ/dq (3 x (N-1) col-major)
    // and dJ/dT (N).  Upstream spelling kept: propogateGrad.
    void propogateGrad(const double *partialGradByCoeffs, const double *partialGradByTimes,
                       double *gradByPoints, double *gradByTimes) const {
        std::vector<double> adj(partialGradByCoeffs, partialGradByCoeffs + static_cast<size_t>(D) * N * 3);
        A.solveAdj(adj.data(), 3);
        for (int i = 0; i < N - 1; ++i)
            for (int a = 0; a < 3; ++a) gradByPoints[i * 3 + a] = adj[(D * i + D - 1) * 3 + a];
        double beta[D];
        for (int i = 0; i < N; ++i) {
            double acc = 0.0;
            auto row_term = [&](int row, int d) {
                // adj.row(row) . ( beta^(d+1)(T_i) c_i )
                basis<D>(T[i], d + 1, beta);
                for (int a = 0; a < 3; ++a) {
                    double v = 0.0;
                    for (int k = 0; k < D; ++k) v += beta[k] * b[(D * i + k) * 3 + a];
                    acc += adj[row * 3 + a] * v;
                }
            };
            if (i < N - 1) {
                const int r = D * i + S;
                for (int j = 0; j <= S - 2; ++j) row_term(r + j, S + j);
                row_term(r + S - 1, 0);
                for (int d = 0; d < S; ++d) row_term(r + S + d, d);
            } else {
                for (int d = 0; d < S; ++d) row_term(D * N - S + d, d);
            }
            gradByTimes[i] = partialGradByTimes[i] - acc;
        }
    }

    // Trajectory<2S-1> packing: out[i][axis][k], k = 0 is the HIGHEST power
    // (gcopter/trajectory.hpp:79-83; flatten index of planner/qp_solver.hpp:133
    //  idx = i*3*d + j*d + k consumed at planner/learning_planner.hpp:212).
    void getTrajectoryFlat(double *out) const {
        for (int i = 0; i < N; ++i)
            for (int a = 0; a < 3; ++a)
                for (int k = 0; k < D; ++k)
                    out[(i * 3 + a) * D + k] = b[(D * i + (D - 1 - k)) * 3 + a];
    }

private:
    static double fallfac(int a) {  // a!/(a-S)!
        double f = 1.0;
        for (int u = 0; u < S; ++u) f *= static_cast<double>(a - u);
        return f;
    }
    static double dot3(const double *u, const double *v) { return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]; }

    int N = 0;
    double head[3 * S], tail[3 * S];
    Banded A;
    std::vector<double> b, T;
};

// ---------------------------------------------------------------------------
// smoothed hinge, restating gcopter/firi.hpp:60-84 (smoothedL1).
// ---------------------------------------------------------------------------
inline bool smoothed_l1(double mu, double x, double &f, double &df) {
    if (x < 0.0) return false;
    if (x > mu) { f = x - 0.5 * mu; df = 1.0; return true; }
    const double r = x / mu, r2 = r * r, h = mu - 0.5 * x;
    f = h * r2 * r;
    df = r2 * (-0.5 * r + 3.0 * h / mu);
    return true;
}

// tau <-> T diffeomorphism (upstream gcopter.hpp forwardT/backwardT/backwardGradT).
inline double forward_t(double tau) {
    return tau > 0.0 ? ((0.5 * tau + 1.0) * tau + 1.0) : 1.0 / ((0.5 * tau - 1.0) * tau + 1.0);
}
inline double backward_t(double T) {
    return T > 1.0 ? (std::sqrt(2.0 * T - 1.0) - 1.0) : (1.0 - std::sqrt(2.0 / T - 1.0));
}
inline double backward_grad_t(double tau, double gradT) {
    if (tau > 0.0) return gradT * (tau + 1.0);
    const double den = (0.5 * tau - 1.0) * tau + 1.0;
    return gradT * (1.0 - tau) / (den * den);
}

struct PenaltyParams {
    int kappa = 16;        // IntegralIntervs
    double mu = 1.0e-2;    // SmoothingEps (config/planner.yaml:15)
    double w_pos = 1.0e4, w_vel = 1.0e4, w_acc = 1.0e4, w_jerk = 1.0e4;
    double v_max = 4.0, a_max = 6.0, j_max = 12.0;
    double rho = 20.0;     // WeightT
    // fixed-time mode: T is data (the call the reference makes today, learning_planner.hpp:196 -> qp_solver.hpp:119):
    // the tau block of the gradient is zero, so lbfgs_optimize never moves the durations
    bool freeze_times = false;
    // rows are [n, b] with n.p <= b (after the sign flip of learning_planner.hpp:293-299) instead of n.p + d <= 0
    bool planner_rows = false;
};

// One corridor problem.  hpolys: [N][Kstride][4] rows (nx,ny,nz,d), GCOPTER sign
// n.p + d <= 0 (gcopter/geo_utils.hpp:41-42); hrows[i] <= Kstride rows used.
struct Problem {
    int N = 0;
    const double *head = nullptr, *tail = nullptr;  // 3 x S col-major
    const double *hpolys = nullptr;
    const int *hrows = nullptr;
    int Kstride = 0;
};

// Penalty sampling along every piece.
template <int S>
inline void attach_penalty(const PenaltyParams &pp, const Problem &pb, const double *T,
                           const double *coeffs, double &cost, double *gdT, double *gdC) {
    constexpr int D = 2 * S;
    const int N = pb.N, kap = pp.kappa;
    double bb[5][D];
    double (&b0)[D] = bb[0], (&b1)[D] = bb[1], (&b2)[D] = bb[2], (&b3)[D] = bb[3], (&b4)[D] = bb[4];
    const double vmax2 = pp.v_max * pp.v_max, amax2 = pp.a_max * pp.a_max, jmax2 = pp.j_max * pp.j_max;
    for (int i = 0; i < N; ++i) {
        const double *c = coeffs + static_cast<size_t>(D) * i * 3;
        const double step = T[i] / kap;
        const int K = pb.hrows ? pb.hrows[i] : 0;
        const double *hp = pb.hpolys ? pb.hpolys + static_cast<size_t>(i) * pb.Kstride * 4 : nullptr;
        for (int j = 0; j <= kap; ++j) {
            const double s = j * step;
            basis_rows<D, 5>(s, bb);
            double pos[3] = {0, 0, 0}, vel[3] = {0, 0, 0}, acc[3] = {0, 0, 0}, jer[3] = {0, 0, 0}, sna[3] = {0, 0, 0};
            for (int k = 0; k < D; ++k)
                for (int a = 0; a < 3; ++a) {
                    const double ck = c[k * 3 + a];
                    pos[a] += b0[k] * ck; vel[a] += b1[k] * ck; acc[a] += b2[k] * ck;
                    jer[a] += b3[k] * ck; sna[a] += b4[k] * ck;
                }
            const double node = (j == 0 || j == kap) ? 0.5 : 1.0;
            const double alpha = static_cast<double>(j) / kap;
            double pena = 0.0, gP[3] = {0, 0, 0}, gV[3] = {0, 0, 0}, gA[3] = {0, 0, 0}, gJ[3] = {0, 0, 0};
            double f, df;
            bool any_active = false;
            for (int k = 0; k < K; ++k) {
                const double *h = hp + k * 4;
                const double viol = h[0] * pos[0] + h[1] * pos[1] + h[2] * pos[2] + (pp.planner_rows ? -h[3] : h[3]);
                if (smoothed_l1(pp.mu, viol, f, df)) {
                    any_active = true;
                    for (int a = 0; a < 3; ++a) gP[a] += pp.w_pos * df * h[a];
                    pena += pp.w_pos * f;
                }
            }
            const double vv = vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2] - vmax2;
            if (smoothed_l1(pp.mu, vv, f, df)) {
                any_active = true;
                for (int a = 0; a < 3; ++a) gV[a] += pp.w_vel * df * 2.0 * vel[a];
                pena += pp.w_vel * f;
            }
            const double aa = acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2] - amax2;
            if (smoothed_l1(pp.mu, aa, f, df)) {
                any_active = true;
                for (int a = 0; a < 3; ++a) gA[a] += pp.w_acc * df * 2.0 * acc[a];
                pena += pp.w_acc * f;
            }
            const double jj = jer[0] * jer[0] + jer[1] * jer[1] + jer[2] * jer[2] - jmax2;
            if (smoothed_l1(pp.mu, jj, f, df)) {
                any_active = true;
                for (int a = 0; a < 3; ++a) gJ[a] += pp.w_jerk * df * 2.0 * jer[a];
                pena += pp.w_jerk * f;
            }
            if (!any_active) continue;   // nothing violated at this sample: every term below is an exact zero
            const double w = node * step;
            for (int k = 0; k < D; ++k)
                for (int a = 0; a < 3; ++a)
                    gdC[(D * i + k) * 3 + a] += (b0[k] * gP[a] + b1[k] * gV[a] + b2[k] * gA[a] + b3[k] * gJ[a]) * w;
            double dsum = 0.0;
            for (int a = 0; a < 3; ++a) dsum += gP[a] * vel[a] + gV[a] * acc[a] + gA[a] * jer[a] + gJ[a] * sna[a];
            gdT[i] += dsum * alpha * w + node * pena / kap;
            cost += w * pena;
        }
    }
}

// costFunctional: x = [tau(N); q(3(N-1)) as [i*3+axis]] -> f, g (same layout).
template <int S>
struct CostFunctional {
    static constexpr int D = 2 * S;
    PenaltyParams pp;
    Problem pb;
    Minco<S> minco;
    std::vector<double> T, gdC, gdT, gq, gT;
    long evals = 0;

    void setup(const PenaltyParams &p, const Problem &prob) {
        pp = p; pb = prob;
        minco.setConditions(pb.head, pb.tail, pb.N);
        T.resize(pb.N); gdT.resize(pb.N); gT.resize(pb.N);
        gdC.resize(static_cast<size_t>(D) * pb.N * 3);
        gq.resize(static_cast<size_t>(3) * std::max(pb.N - 1, 1));
    }
    int nvars() const { return pb.N + 3 * (pb.N - 1); }

    double eval(const double *x, double *g) {
        const int N = pb.N;
        ++evals;
        for (int i = 0; i < N; ++i) T[i] = forward_t(x[i]);
        minco.setParameters(x + N, T.data());
        double cost;
        minco.getEnergy(cost);
        minco.getEnergyPartialGradByCoeffs(gdC.data());
        minco.getEnergyPartialGradByTimes(gdT.data());
        attach_penalty<S>(pp, pb, T.data(), minco.getCoeffs().data(), cost, gdT.data(), gdC.data());
        minco.propogateGrad(gdC.data(), gdT.data(), gq.data(), gT.data());
        double tsum = 0.0;
        for (int i = 0; i < N; ++i) tsum += T[i];
        cost += pp.rho * tsum;
        for (int i = 0; i < N; ++i) g[i] = pp.freeze_times ? 0.0 : backward_grad_t(x[i], gT[i] + pp.rho);
        for (int i = 0; i < 3 * (N - 1); ++i) g[N + i] = gq[i];
        return cost;
    }
    static double thunk(void *self, const double *x, double *g, int) {
        return static_cast<CostFunctional *>(self)->eval(x, g);
    }
};

}  // namespace orc
