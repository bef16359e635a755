// launch.h -- host-visible launch table; one instance per (S, LPT) object (kernels_inst.cu).
#pragma once
#include <cuda_runtime.h>

#include "args.h"

namespace mincob {
struct LaunchResult {
    cudaError_t err;
    int code;      // 0 or MINCOB_E_INVALID when the kernel cannot be resident
    size_t smem;
    int mapping;   // optimize: MINCOB_MAP_THROUGHPUT / MINCOB_MAP_LATENCY actually launched
};
struct LaunchTable {
    LaunchResult (*evaluate)(cudaStream_t, int sm_count, const DevParams &, const BatchArgs &);
    LaunchResult (*optimize)(cudaStream_t, int sm_count, const DevParams &, const BatchArgs &);
    LaunchResult (*minco)(cudaStream_t, int sm_count, const MincoArgs &, int propagate);
    // bytes of global scratch (L-BFGS history slabs) `optimize` needs in BatchArgs::hist
    size_t (*optimize_scratch)(int sm_count, const DevParams &, const BatchArgs &);
    LaunchResult (*check)(cudaStream_t, int sm_count, const CheckArgs &);
    LaunchResult (*maxrates)(cudaStream_t, int sm_count, const RateArgs &);
};
}  // namespace mincob
const mincob::LaunchTable *mincob_table_3_5();
const mincob::LaunchTable *mincob_table_4_5();
const mincob::LaunchTable *mincob_table_3_8();
const mincob::LaunchTable *mincob_table_3_16();
const mincob::LaunchTable *mincob_table_3_32();
const mincob::LaunchTable *mincob_table_4_8();
const mincob::LaunchTable *mincob_table_4_16();
const mincob::LaunchTable *mincob_table_4_32();
