// ============================================================================
// kernels_inst.cu -- one (S, LPT) instantiation of every kernel plus its launchers.
// Built once per pair with -DMINCOB_S=.. -DMINCOB_LPT=.. (allocnet_b200/build.py) so the six
// objects compile in parallel.
// ============================================================================
#include "../../include/mincob.h"
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
    const int lig = (threadIdx.x & 31) % LPT;
    const unsigned mask = group_mask<LPT>();
    const int N = a.N;
    const int groups = gridDim.x * (THREADS / LPT);
    const int rounds = (a.B + groups - 1) / groups;
    int p = blockIdx.x * (THREADS / LPT) + threadIdx.x / LPT;
    for (int it = 0; it < rounds; ++it, p += groups) {
        const bool live = p < a.B;
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
        Spline<S, LPT> sp;
        double chat[D][3];
        spline_solve<S, LPT>(mask, lig, Ne, T, P0, P1, hd, td, sp, chat);
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
            spline_adjoint<S, LPT>(mask, lig, Ne, sp, G, gTp, gq, gT);
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

}  // namespace mincob

namespace {
using namespace mincob;
constexpr int S = MINCOB_S, LPT = MINCOB_LPT, THREADS = 128, GPB = THREADS / LPT;

LaunchResult ok(cudaError_t e) { return LaunchResult{e, 0, 0}; }

LaunchResult launch_evaluate(cudaStream_t st, int sm_count, const DevParams &dp, const BatchArgs &a) {
    int blocks = (a.B + GPB - 1) / GPB;
    const int cap = sm_count * 16;
    if (blocks > cap) blocks = cap;
    evaluate_kernel<S, LPT, THREADS><<<blocks, THREADS, 0, st>>>(dp, a);
    return ok(cudaGetLastError());
}

LaunchResult launch_optimize(cudaStream_t st, int sm_count, const DevParams &dp, const BatchArgs &a) {
    const int m = dp.mem, past = dp.past > 0 ? dp.past : 1;
    const size_t smem = (size_t)GPB * (2 * m * 4 * LPT + 2 * m + past) * sizeof(double);
    auto kern = optimize_kernel<S, LPT, THREADS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return LaunchResult{cudaSuccess, MINCOB_E_INVALID, smem};
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (e != cudaSuccess) return ok(e);
    if (per_sm < 1) return LaunchResult{cudaSuccess, MINCOB_E_INVALID, smem};
    int blocks = per_sm * sm_count;
    const int need = (a.B + GPB - 1) / GPB;
    if (blocks > need) blocks = need;
    if ((e = cudaMemsetAsync(a.counter, 0, sizeof(int), st)) != cudaSuccess) return ok(e);
    if ((e = cudaMemsetAsync(a.total_evals, 0, sizeof(unsigned long long), st)) != cudaSuccess) return ok(e);
    kern<<<blocks, THREADS, smem, st>>>(dp, a);
    return ok(cudaGetLastError());
}

LaunchResult launch_minco(cudaStream_t st, int sm_count, const MincoArgs &a, int propagate) {
    int blocks = (a.B + GPB - 1) / GPB;
    const int cap = sm_count * 16;
    if (blocks > cap) blocks = cap;
    minco_kernel<S, LPT, THREADS><<<blocks, THREADS, 0, st>>>(a, propagate);
    return ok(cudaGetLastError());
}

const LaunchTable kTable = {launch_evaluate, launch_optimize, launch_minco};
}  // namespace

#define MINCOB_CAT_(a, b, c) mincob_table_##a##_##b
#define MINCOB_CAT(a, b) MINCOB_CAT_(a, b, )
const mincob::LaunchTable *MINCOB_CAT(MINCOB_S, MINCOB_LPT)() { return &kTable; }
