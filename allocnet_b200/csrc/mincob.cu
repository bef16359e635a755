// ============================================================================
// mincob.cu -- C-ABI (include/mincob.h) over the sm_100a kernels.  Host side is plain C++:
// argument checks, device buffers, launches, stream/event plumbing, NCCL via dlopen.
// There is no CPU compute path in this file: without a CUDA device every compute entry point
// fails with MINCOB_E_CUDA.
// ============================================================================
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/mincob.h"
#include "launch.h"

using namespace mincob;

// ---- NCCL through dlopen (so the library loads, and its symbols can be listed, without NCCL) --
typedef struct ncclComm *ncclComm_t;
struct NcclUid { char b[128]; };
typedef int (*nccl_get_uid_t)(NcclUid *);
typedef int (*nccl_init_rank_t)(ncclComm_t *, int, NcclUid, int);
typedef int (*nccl_allgather_t)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
typedef int (*nccl_destroy_t)(ncclComm_t);
typedef const char *(*nccl_errstr_t)(int);

static struct {
    void *lib;
    nccl_get_uid_t get_uid;
    nccl_init_rank_t init_rank;
    nccl_allgather_t allgather;
    nccl_destroy_t destroy;
    nccl_errstr_t errstr;
} g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

static bool nccl_load() {
    if (g_nccl.lib) return true;
    // RTLD_DEFAULT first: inside a torch process libnccl is already mapped.
    void *probe = dlsym(RTLD_DEFAULT, "ncclGetUniqueId");
    void *lib = nullptr;
    if (!probe) {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
    }
    void *h = lib ? lib : RTLD_DEFAULT;
    g_nccl.get_uid = (nccl_get_uid_t)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (nccl_init_rank_t)dlsym(h, "ncclCommInitRank");
    g_nccl.allgather = (nccl_allgather_t)dlsym(h, "ncclAllGather");
    g_nccl.destroy = (nccl_destroy_t)dlsym(h, "ncclCommDestroy");
    g_nccl.errstr = (nccl_errstr_t)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allgather || !g_nccl.destroy) return false;
    g_nccl.lib = lib ? lib : (void *)1;
    return true;
}

// ---- handle -------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct mincob_ctx {
    mincob_params prm;
    DevParams dp;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr, ev_copy0 = nullptr;
    int *ready = nullptr;        // device: problems uploaded so far (mincob_set_problems_async)
    int *ready_host = nullptr;   // pinned: the values written to `ready`, one per chunk
    bool upload_in_flight = false;
    // mincob_set_problems_async only records the request; the copies are enqueued by the next call that needs the data,
    // so that an optimize call can put its own (small) x upload in front of them in the copy queue
    bool upload_pending = false;
    const double *p_head = nullptr, *p_tail = nullptr, *p_hpolys = nullptr;
    const int32_t *p_hrows = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int last_launches = 0, last_mapping = 0;
    bool timed = false;
    int sm_count = 0;
    // problems (device pointers; owned only when they point into the DevBufs below)
    int B = 0, N = 0, K = 0;
    const double *head = nullptr, *tail = nullptr, *hpolys = nullptr;
    const int *hrows = nullptr;
    DevBuf b_head, b_tail, b_hpolys, b_hrows;          // set_problems (host) staging
    DevBuf b_hist;                                     // L-BFGS (s, y) history slabs of the resident groups
    DevBuf b_x, b_f, b_g, b_status, b_iters, b_evals, b_coeffs, b_T, b_gather;  // host-pointer entry points
    DevBuf b_m0, b_m1, b_m2, b_m3, b_m4, b_m5, b_m6, b_m7, b_m8;      // minco_forward / propagate
    int *counter = nullptr;
    unsigned long long *total_evals = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    int64_t gathered_count = 0;   // doubles in b_gather after the last sharded optimize call
    std::string err;
};

static int fail(mincob_ctx *h, int code, const char *fmt, ...) {
    if (h) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        h->err = buf;
    }
    return code;
}
#define CU(h, call)                                                                                \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(h, MINCOB_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// every entry point runs on the handle's device and restores the caller's current device on return
struct DeviceGuard {
    int prev = -1;
    bool ok;
    explicit DeviceGuard(int dev) {
        ok = cudaGetDevice(&prev) == cudaSuccess;
        if (ok && prev != dev) ok = cudaSetDevice(dev) == cudaSuccess; else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(h)                \
    DeviceGuard guard_(h->device);  \
    if (!guard_.ok) return fail(h, MINCOB_E_CUDA, "cudaSetDevice(%d) failed", h->device)

static int ensure(mincob_ctx *h, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    if (bytes == 0) bytes = 8;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) return fail(h, MINCOB_E_ALLOC, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    b.cap = bytes;
    return 0;
}
static void release(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

static int derive(mincob_ctx *h) {
    const mincob_params &p = h->prm;
    if (p.S != 3 && p.S != 4) return fail(h, MINCOB_E_INVALID, "S must be 3 or 4 (got %d)", p.S);
    if (p.kappa < 1) return fail(h, MINCOB_E_INVALID, "kappa must be >= 1");
    if (!(p.mu > 0.0)) return fail(h, MINCOB_E_INVALID, "mu must be > 0");
    DevParams &d = h->dp;
    d.kappa = p.kappa; d.mu = p.mu;
    d.imu = 1.0 / p.mu; d.ikap = 1.0 / p.kappa;
    d.w_pos = p.w_pos; d.w_vel = p.w_vel; d.w_acc = p.w_acc; d.w_jerk = p.w_jerk;
    d.vmax2 = p.v_max * p.v_max; d.amax2 = p.a_max * p.a_max; d.jmax2 = p.j_max * p.j_max;
    d.amax2q = 0.25 * d.amax2;
    d.rho = p.rho;
    if (p.flags & ~(MINCOB_FLAG_FREEZE_TIMES | MINCOB_FLAG_PLANNER_ROWS)) return fail(h, MINCOB_E_INVALID, "unknown bits in flags (%d)", p.flags);
    if (p.mapping < MINCOB_MAP_AUTO || p.mapping > MINCOB_MAP_LATENCY) return fail(h, MINCOB_E_INVALID, "mapping must be MINCOB_MAP_AUTO/THROUGHPUT/LATENCY");
    d.freeze = (p.flags & MINCOB_FLAG_FREEZE_TIMES) ? 1 : 0;
    d.mapping = p.mapping;
    d.penalties = (p.w_pos != 0.0 || p.w_vel != 0.0 || p.w_acc != 0.0 || p.w_jerk != 0.0) ? 1 : 0;
    d.mem = p.mem_size; d.past = p.past; d.max_iter = p.max_iterations; d.max_ls = p.max_linesearch;
    d.g_eps = p.g_epsilon; d.delta = p.delta; d.min_step = p.min_step; d.max_step = p.max_step;
    d.f_dec = p.f_dec_coeff; d.s_curv = p.s_curv_coeff; d.cautious = p.cautious_factor; d.mach_prec = p.machine_prec;
    return 0;
}

// lbfgs.hpp:450-495 parameter validation; returns 0 or the LBFGSERR_* the reference would return.
static int lbfgs_param_code(const mincob_params &p, int n) {
    if (n <= 0) return LBFGSERR_INVALID_N;
    if (p.mem_size <= 0) return LBFGSERR_INVALID_MEMSIZE;
    if (p.g_epsilon < 0.0) return LBFGSERR_INVALID_GEPSILON;
    if (p.past < 0) return LBFGSERR_INVALID_TESTPERIOD;
    if (p.delta < 0.0) return LBFGSERR_INVALID_DELTA;
    if (p.min_step < 0.0) return LBFGSERR_INVALID_MINSTEP;
    if (p.max_step < p.min_step) return LBFGSERR_INVALID_MAXSTEP;
    if (!(p.f_dec_coeff > 0.0 && p.f_dec_coeff < 1.0)) return LBFGSERR_INVALID_FDECCOEFF;
    if (!(p.s_curv_coeff < 1.0 && p.s_curv_coeff > p.f_dec_coeff)) return LBFGSERR_INVALID_SCURVCOEFF;
    if (!(p.machine_prec > 0.0)) return LBFGSERR_INVALID_MACHINEPREC;
    if (p.max_linesearch <= 0) return LBFGSERR_INVALID_MAXLINESEARCH;
    return 0;
}

// kernels are instantiated per (S, LPT) in kernels_inst.cu (one object each, built in parallel)
// lanes per trajectory: 5 (six trajectories per warp; the reference plans with at most ModelMaxSeg = 5 pieces,
// learning_planner.hpp:33), 8, 16 or 32
static int lpt_index(int N) { return N <= 5 ? 0 : (N <= 8 ? 1 : (N <= 16 ? 2 : 3)); }
static const LaunchTable *table_for(int S, int N) {
    static const LaunchTable *tabs[2][4] = {
        {mincob_table_3_5(), mincob_table_3_8(), mincob_table_3_16(), mincob_table_3_32()},
        {mincob_table_4_5(), mincob_table_4_8(), mincob_table_4_16(), mincob_table_4_32()}};
    return tabs[S == 3 ? 0 : 1][lpt_index(N)];
}
static int launched(mincob_ctx *h, const LaunchResult &r, const char *what) {
    if (r.err != cudaSuccess) return fail(h, MINCOB_E_CUDA, "%s: %s", what, cudaGetErrorString(r.err));
    if (r.code) return fail(h, r.code, "%s: kernel does not fit (dynamic smem %zu B); lower mem_size", what, r.smem);
    return 0;
}
static int do_evaluate(mincob_ctx *h, const BatchArgs &a) {
    return launched(h, table_for(h->prm.S, a.N)->evaluate(h->stream, h->sm_count, h->dp, a), "evaluate_kernel");
}
static int do_optimize(mincob_ctx *h, BatchArgs &a) {
    const LaunchTable *t = table_for(h->prm.S, a.N);
    const size_t need = t->optimize_scratch(h->sm_count, h->dp, a);
    const void *before = h->b_hist.p;
    int rc = ensure(h, h->b_hist, need);
    if (rc) return rc;
    // The two-loop recursion reads history slots a problem has not written yet (their contribution is predicated
    // off): give a fresh slab defined contents once, so those reads never see arbitrary bit patterns.
    if (h->b_hist.p != before) CU(h, cudaMemsetAsync(h->b_hist.p, 0, h->b_hist.cap, h->stream));
    a.hist = (double *)h->b_hist.p;
    const LaunchResult r = t->optimize(h->stream, h->sm_count, h->dp, a);
    h->last_mapping = r.mapping;
    return launched(h, r, "optimize_kernel");
}
static int do_minco(mincob_ctx *h, const MincoArgs &a, int propagate) {
    return launched(h, table_for(h->prm.S, a.N)->minco(h->stream, h->sm_count, a, propagate), "minco_kernel");
}

// Enqueue a recorded mincob_set_problems_async request on the copy stream: first `x_host` -> `x_dev` when the caller is a
// host-pointer optimize call (its start vector must arrive BEFORE the bulk of the batch, the copy queue is first in first
// out), then the arrival counter reset, then the batch in problem order, a counter update after every chunk.
static int start_upload(mincob_ctx *h, const double *x_host = nullptr, double *x_dev = nullptr, size_t x_bytes = 0) {
    if (!h->upload_pending) return 0;
    h->upload_pending = false;
    const int B = h->B, N = h->N, K = h->K, S = h->prm.S;
    const size_t pb = (size_t)S * 3 * sizeof(double), pp = (size_t)N * K * 4 * sizeof(double), pr = (size_t)N * sizeof(int);
    // the copy stream starts after whatever the compute stream has queued so far (a previous kernel may still read the buffers)
    CU(h, cudaEventRecord(h->ev_copy, h->stream));
    CU(h, cudaStreamWaitEvent(h->copy_stream, h->ev_copy, 0));
    if (x_host) CU(h, cudaMemcpyAsync(x_dev, x_host, x_bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CU(h, cudaMemsetAsync(h->ready, 0, sizeof(int), h->copy_stream));
    CU(h, cudaEventRecord(h->ev_copy0, h->copy_stream));      // the optimize kernel must not start before x is there and the counter is reset
    // chunk = a multiple of 4096 problems: no 128-byte line of head / tail / hrows / hpolys straddles two chunks, so a line
    // an SM has cached never holds problems that had not arrived when it was read
    int per = 4096;
    while ((B + per - 1) / per > 4096) per *= 2;
    for (int c = 0, lo = 0; lo < B; ++c, lo += per) {
        const int cnt = (lo + per <= B) ? per : B - lo;
        CU(h, cudaMemcpyAsync((char *)h->b_head.p + pb * lo, (const char *)h->p_head + pb * lo, pb * cnt, cudaMemcpyHostToDevice, h->copy_stream));
        CU(h, cudaMemcpyAsync((char *)h->b_tail.p + pb * lo, (const char *)h->p_tail + pb * lo, pb * cnt, cudaMemcpyHostToDevice, h->copy_stream));
        if (K > 0) {
            CU(h, cudaMemcpyAsync((char *)h->b_hpolys.p + pp * lo, (const char *)h->p_hpolys + pp * lo, pp * cnt, cudaMemcpyHostToDevice, h->copy_stream));
            CU(h, cudaMemcpyAsync((char *)h->b_hrows.p + pr * lo, (const char *)h->p_hrows + pr * lo, pr * cnt, cudaMemcpyHostToDevice, h->copy_stream));
        }
        h->ready_host[c] = lo + cnt;
        CU(h, cudaMemcpyAsync(h->ready, h->ready_host + c, sizeof(int), cudaMemcpyHostToDevice, h->copy_stream));
    }
    CU(h, cudaEventRecord(h->ev_copy, h->copy_stream));
    h->upload_in_flight = true;
    return 0;
}

// every consumer of the problem buffers other than the optimize kernel waits for the whole upload
static int upload_done(mincob_ctx *h) {
    int rc = start_upload(h);
    if (rc) return rc;
    if (h->upload_in_flight) {
        CU(h, cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
        h->upload_in_flight = false;
    }
    return 0;
}
// may_overlap: the caller (the optimize kernel) follows a chunked upload problem by problem; everybody else waits for it
static int have_problems(mincob_ctx *h, bool may_overlap = false) {
    if (!h) return MINCOB_E_INVALID;
    if (h->B <= 0 || !h->head || !h->tail) return fail(h, MINCOB_E_STATE, "set_problems has not been called");
    return may_overlap ? 0 : upload_done(h);
}
static BatchArgs base_args(mincob_ctx *h) {
    BatchArgs a;
    memset(&a, 0, sizeof a);
    a.B = h->B; a.N = h->N; a.K = h->K;
    a.head = h->head; a.tail = h->tail; a.hpolys = h->hpolys; a.hrows = h->hrows;
    a.counter = h->counter; a.total_evals = h->total_evals;
    a.ready = nullptr;
    return a;
}

extern "C" {

int mincob_version(void) { return 100; }

int mincob_default_params(mincob_params *p, int S) {
    if (!p || (S != 3 && S != 4)) return MINCOB_E_INVALID;
    memset(p, 0, sizeof *p);
    p->S = S; p->kappa = 16; p->mu = 1.0e-2;
    p->w_pos = p->w_vel = p->w_acc = p->w_jerk = 1.0e4;
    p->v_max = 4.0; p->a_max = 6.0; p->j_max = 12.0; p->rho = 20.0;
    p->mem_size = 8; p->past = 3; p->max_iterations = 1000; p->max_linesearch = 64;
    p->g_epsilon = 0.0; p->delta = 1.0e-5; p->min_step = 1.0e-32; p->max_step = 1.0e20;
    p->f_dec_coeff = 1.0e-4; p->s_curv_coeff = 0.9; p->cautious_factor = 1.0e-6; p->machine_prec = 1.0e-16;
    p->flags = 0; p->mapping = MINCOB_MAP_AUTO;
    return 0;
}

const char *mincob_strerror(int code) {
    switch (code) {
        case MINCOB_OK: return "ok";
        case MINCOB_E_INVALID: return "invalid argument";
        case MINCOB_E_CUDA: return "CUDA error or no CUDA device (this library has no CPU path)";
        case MINCOB_E_STATE: return "call order error";
        case MINCOB_E_NCCL: return "NCCL unavailable or failed";
        case MINCOB_E_ALLOC: return "device allocation failed";
        default: return "unknown mincob error";
    }
}

// Messages follow lbfgs_strerror, gcopter/lbfgs.hpp:724-800 (same meaning per code).
const char *mincob_lbfgs_strerror(int s) {
    switch (s) {
        case LBFGS_CONVERGENCE: return "Success: reached convergence (g_epsilon).";
        case LBFGS_STOP: return "Success: met stopping criteria (past f decrease less than delta).";
        case LBFGS_CANCELED: return "The iteration has been canceled by the monitor callback.";
        case LBFGSERR_UNKNOWNERROR: return "Unknown error.";
        case LBFGSERR_INVALID_N: return "Invalid number of variables specified.";
        case LBFGSERR_INVALID_MEMSIZE: return "Invalid parameter lbfgs_parameter_t::mem_size specified.";
        case LBFGSERR_INVALID_GEPSILON: return "Invalid parameter lbfgs_parameter_t::g_epsilon specified.";
        case LBFGSERR_INVALID_TESTPERIOD: return "Invalid parameter lbfgs_parameter_t::past specified.";
        case LBFGSERR_INVALID_DELTA: return "Invalid parameter lbfgs_parameter_t::delta specified.";
        case LBFGSERR_INVALID_MINSTEP: return "Invalid parameter lbfgs_parameter_t::min_step specified.";
        case LBFGSERR_INVALID_MAXSTEP: return "Invalid parameter lbfgs_parameter_t::max_step specified.";
        case LBFGSERR_INVALID_FDECCOEFF: return "Invalid parameter lbfgs_parameter_t::f_dec_coeff specified.";
        case LBFGSERR_INVALID_SCURVCOEFF: return "Invalid parameter lbfgs_parameter_t::s_curv_coeff specified.";
        case LBFGSERR_INVALID_MACHINEPREC: return "Invalid parameter lbfgs_parameter_t::machine_prec specified.";
        case LBFGSERR_INVALID_MAXLINESEARCH: return "Invalid parameter lbfgs_parameter_t::max_linesearch specified.";
        case LBFGSERR_INVALID_FUNCVAL: return "The function value became NaN or Inf.";
        case LBFGSERR_MINIMUMSTEP: return "The line-search step became smaller than lbfgs_parameter_t::min_step.";
        case LBFGSERR_MAXIMUMSTEP: return "The line-search step became larger than lbfgs_parameter_t::max_step.";
        case LBFGSERR_MAXIMUMLINESEARCH: return "Line search reaches the maximum try number, assumptions not satisfied or precision not achievable.";
        case LBFGSERR_MAXIMUMITERATION: return "The algorithm routine reaches the maximum number of iterations.";
        case LBFGSERR_WIDTHTOOSMALL: return "Relative search interval width is at least lbfgs_parameter_t::machine_prec.";
        case LBFGSERR_INVALIDPARAMETERS: return "A logic error (negative line-search step) occurred.";
        case LBFGSERR_INCREASEGRADIENT: return "The current search direction increases the cost function value.";
        default: return "(unknown)";
    }
}

const char *mincob_last_error(mincob_handle h) { return h ? h->err.c_str() : "null handle"; }

int mincob_create(mincob_handle *out, const mincob_params *params, int device) {
    if (!out || !params) return MINCOB_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return MINCOB_E_CUDA;
    mincob_ctx *h = new (std::nothrow) mincob_ctx;
    if (!h) return MINCOB_E_ALLOC;
    h->prm = *params;
    h->device = device;
    int rc = derive(h);
    if (rc) { delete h; return rc; }
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return MINCOB_E_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return MINCOB_E_CUDA; }
    h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return MINCOB_E_CUDA; }
    h->stream = h->own_stream;
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_copy0, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&h->ready, sizeof(int)) != cudaSuccess ||
        cudaHostAlloc((void **)&h->ready_host, 4096 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
        mincob_destroy(h);
        return MINCOB_E_ALLOC;
    }
    if (cudaMalloc((void **)&h->counter, sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&h->total_evals, 128 * sizeof(unsigned long long)) != cudaSuccess) {
        mincob_destroy(h);
        return MINCOB_E_ALLOC;
    }
    *out = h;
    return 0;
}

int mincob_destroy(mincob_handle h) {
    if (!h) return MINCOB_E_INVALID;
    cudaSetDevice(h->device);
    if (h->comm && g_nccl.destroy) g_nccl.destroy(h->comm);
    DevBuf *bufs[] = {&h->b_head, &h->b_tail, &h->b_hpolys, &h->b_hist, &h->b_hrows, &h->b_x, &h->b_f, &h->b_g, &h->b_status,
                      &h->b_iters, &h->b_evals, &h->b_coeffs, &h->b_T, &h->b_gather, &h->b_m0, &h->b_m1, &h->b_m2, &h->b_m3,
                      &h->b_m4, &h->b_m5, &h->b_m6, &h->b_m7, &h->b_m8};
    for (DevBuf *b : bufs) release(*b);
    if (h->counter) cudaFree(h->counter);
    if (h->total_evals) cudaFree(h->total_evals);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->ev_copy0) cudaEventDestroy(h->ev_copy0);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ready) cudaFree(h->ready);
    if (h->ready_host) cudaFreeHost(h->ready_host);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return 0;
}

int mincob_set_params(mincob_handle h, const mincob_params *p) {
    if (!h || !p) return MINCOB_E_INVALID;
    const mincob_params old = h->prm;
    h->prm = *p;
    int rc = derive(h);
    if (rc) { h->prm = old; derive(h); }
    return rc;
}

int mincob_set_stream(mincob_handle h, void *s) {
    if (!h) return MINCOB_E_INVALID;
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return 0;
}

int mincob_synchronize(mincob_handle h) {
    if (!h) return MINCOB_E_INVALID;
    ON_DEVICE(h);
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mincob_last_mapping(mincob_handle h, int *mapping) {
    if (!h || !mapping) return MINCOB_E_INVALID;
    if (!h->last_mapping) return fail(h, MINCOB_E_STATE, "no optimize call has been launched yet");
    *mapping = h->last_mapping;
    return 0;
}

int mincob_last_kernel_ms(mincob_handle h, float *ms, int *launches) {
    if (!h) return MINCOB_E_INVALID;
    if (!h->timed) return fail(h, MINCOB_E_STATE, "no evaluate/optimize call has been timed yet");
    CU(h, cudaEventSynchronize(h->ev1));
    float t = 0.f;
    CU(h, cudaEventElapsedTime(&t, h->ev0, h->ev1));
    if (ms) *ms = t;
    if (launches) *launches = h->last_launches;
    return 0;
}

static int check_shape(mincob_ctx *h, int B, int N, int K) {
    if (B <= 0) return fail(h, MINCOB_E_INVALID, "B must be > 0");
    if (N < 1 || N > MINCOB_MAX_PIECES) return fail(h, MINCOB_E_INVALID, "N must be in [1,%d] (got %d)", MINCOB_MAX_PIECES, N);
    if (K < 0 || K > MINCOB_MAX_ROWS) return fail(h, MINCOB_E_INVALID, "K must be in [0,%d] (got %d)", MINCOB_MAX_ROWS, K);
    return 0;
}

// rows [n, b] (n.p <= b, learning_planner.hpp:293-299) -> [n, d] with n.p + d <= 0 (geo_utils.hpp:41-42): d = -b
static __global__ void planner_rows_kernel(const double *src, double *dst, size_t rows) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) {
        const double4 r = reinterpret_cast<const double4 *>(src)[i];
        reinterpret_cast<double4 *>(dst)[i] = make_double4(r.x, r.y, r.z, -r.w);
    }
}

int mincob_set_problems_device(mincob_handle h, int B, int N, int K, const double *head, const double *tail,
                               const double *hpolys, const int32_t *hrows) {
    if (!h || !head || !tail) return MINCOB_E_INVALID;
    int rc = check_shape(h, B, N, K);
    if (rc) return rc;
    if (K > 0 && (!hpolys || !hrows)) return fail(h, MINCOB_E_INVALID, "K > 0 needs hpolys and hrows");
    if (hpolys && ((uintptr_t)hpolys & 31u)) return fail(h, MINCOB_E_INVALID, "hpolys must be 32-byte aligned");
    if ((rc = upload_done(h))) return rc;
    if (K > 0 && (h->prm.flags & MINCOB_FLAG_PLANNER_ROWS)) {
        // the kernels read GCOPTER-sign rows: keep a converted copy (in place when the rows already are our staging copy)
        ON_DEVICE(h);
        const size_t rows = (size_t)B * N * K;
        double *dst = (double *)h->b_hpolys.p;
        if (hpolys != dst) {
            if ((rc = ensure(h, h->b_hpolys, rows * 4 * sizeof(double)))) return rc;
            dst = (double *)h->b_hpolys.p;
        }
        planner_rows_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, h->stream>>>(hpolys, dst, rows);
        CU(h, cudaGetLastError());
        hpolys = dst;
    }
    h->B = B; h->N = N; h->K = K;
    h->head = head; h->tail = tail;
    h->hpolys = K > 0 ? hpolys : nullptr;
    h->hrows = K > 0 ? (const int *)hrows : nullptr;
    return 0;
}

int mincob_set_problems(mincob_handle h, int B, int N, int K, const double *head, const double *tail,
                        const double *hpolys, const int32_t *hrows) {
    if (!h || !head || !tail) return MINCOB_E_INVALID;
    int rc = check_shape(h, B, N, K);
    if (rc) return rc;
    if (K > 0 && (!hpolys || !hrows)) return fail(h, MINCOB_E_INVALID, "K > 0 needs hpolys and hrows");
    ON_DEVICE(h);
    if ((rc = upload_done(h))) return rc;
    const int S = h->prm.S;
    const size_t nb = (size_t)B * S * 3 * sizeof(double), np = (size_t)B * N * K * 4 * sizeof(double),
                 nr = (size_t)B * N * sizeof(int);
    if ((rc = ensure(h, h->b_head, nb)) || (rc = ensure(h, h->b_tail, nb))) return rc;
    CU(h, cudaMemcpyAsync(h->b_head.p, head, nb, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->b_tail.p, tail, nb, cudaMemcpyHostToDevice, h->stream));
    if (K > 0) {
        if ((rc = ensure(h, h->b_hpolys, np)) || (rc = ensure(h, h->b_hrows, nr))) return rc;
        CU(h, cudaMemcpyAsync(h->b_hpolys.p, hpolys, np, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(h->b_hrows.p, hrows, nr, cudaMemcpyHostToDevice, h->stream));
    }
    rc = mincob_set_problems_device(h, B, N, K, (const double *)h->b_head.p, (const double *)h->b_tail.p,
                                    K > 0 ? (const double *)h->b_hpolys.p : nullptr,
                                    K > 0 ? (const int32_t *)h->b_hrows.p : nullptr);
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->stream));   // host-pointer contract: the caller's buffers are free again on return
    return 0;
}

// Upload in problem order, in chunks, on a second stream; after each chunk a counter in device memory says how many
// problems have arrived.  The next mincob_optimize* call starts at once and its work queue waits per problem for that
// counter, so the kernel overlaps all but the first chunk of the upload.  The caller's buffers must be page-locked
// (mincob_host_alloc / mincob_host_register) and stay untouched until that optimize call has returned (host-pointer
// form) or the stream has been synchronised.
int mincob_set_problems_async(mincob_handle h, int B, int N, int K, const double *head, const double *tail,
                              const double *hpolys, const int32_t *hrows) {
    if (!h || !head || !tail) return MINCOB_E_INVALID;
    int rc = check_shape(h, B, N, K);
    if (rc) return rc;
    if (K > 0 && (!hpolys || !hrows)) return fail(h, MINCOB_E_INVALID, "K > 0 needs hpolys and hrows");
    if (h->prm.flags & MINCOB_FLAG_PLANNER_ROWS) return mincob_set_problems(h, B, N, K, head, tail, hpolys, hrows);   // needs the conversion pass
    ON_DEVICE(h);
    if ((rc = upload_done(h))) return rc;                      // a previous asynchronous batch is uploaded / consumed first
    const int S = h->prm.S;
    const size_t pb = (size_t)S * 3 * sizeof(double), pp = (size_t)N * K * 4 * sizeof(double), pr = (size_t)N * sizeof(int);
    if ((rc = ensure(h, h->b_head, pb * B)) || (rc = ensure(h, h->b_tail, pb * B))) return rc;
    if (K > 0 && ((rc = ensure(h, h->b_hpolys, pp * B)) || (rc = ensure(h, h->b_hrows, pr * B)))) return rc;
    h->p_head = head; h->p_tail = tail; h->p_hpolys = hpolys; h->p_hrows = hrows;
    h->upload_pending = true;
    h->B = B; h->N = N; h->K = K;
    h->head = (const double *)h->b_head.p; h->tail = (const double *)h->b_tail.p;
    h->hpolys = K > 0 ? (const double *)h->b_hpolys.p : nullptr;
    h->hrows = K > 0 ? (const int *)h->b_hrows.p : nullptr;
    return 0;
}

int mincob_evaluate_device(mincob_handle h, const double *x, double *f, double *g) {
    int rc = have_problems(h);
    if (rc) return rc;
    if (!x || !f || !g) return fail(h, MINCOB_E_INVALID, "x, f, g must be non-null");
    ON_DEVICE(h);
    BatchArgs a = base_args(h);
    a.x_in = x; a.f_out = f; a.g_out = g;
    CU(h, cudaEventRecord(h->ev0, h->stream));
    rc = do_evaluate(h, a);
    if (rc) return rc;
    CU(h, cudaEventRecord(h->ev1, h->stream));
    h->timed = true; h->last_launches = 1;
    return 0;
}

int mincob_evaluate(mincob_handle h, const double *x, double *f, double *g) {
    int rc = have_problems(h);
    if (rc) return rc;
    if (!x || !f || !g) return fail(h, MINCOB_E_INVALID, "x, f, g must be non-null");
    ON_DEVICE(h);
    const size_t n = (size_t)h->N + 3 * (h->N - 1), nx = (size_t)h->B * n * sizeof(double), nf = (size_t)h->B * sizeof(double);
    if ((rc = ensure(h, h->b_x, nx)) || (rc = ensure(h, h->b_g, nx)) || (rc = ensure(h, h->b_f, nf))) return rc;
    CU(h, cudaMemcpyAsync(h->b_x.p, x, nx, cudaMemcpyHostToDevice, h->stream));
    rc = mincob_evaluate_device(h, (const double *)h->b_x.p, (double *)h->b_f.p, (double *)h->b_g.p);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(f, h->b_f.p, nf, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(g, h->b_g.p, nx, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

static __global__ void fill_status_kernel(int *status, int B, int code) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) status[i] = code;
}

int mincob_optimize_device(mincob_handle h, double *x, double *f, int32_t *status, int32_t *iters, int32_t *evals,
                           double *coeffs, double *T) {
    int rc = have_problems(h, true);
    if (rc) return rc;
    if (!x) return fail(h, MINCOB_E_INVALID, "x must be non-null");
    ON_DEVICE(h);
    const int n = h->N + 3 * (h->N - 1);
    const int pc = lbfgs_param_code(h->prm, n);
    if (pc) {  // lbfgs.hpp:450-495: the reference returns the code before touching x
        if (status) {
            fill_status_kernel<<<(h->B + 255) / 256, 256, 0, h->stream>>>(status, h->B, pc);
            CU(h, cudaGetLastError());
        }
        h->timed = false;
        return 0;
    }
    if (h->prm.mem_size > MINCOB_MAX_MEM) return fail(h, MINCOB_E_INVALID, "mem_size > %d not supported", MINCOB_MAX_MEM);
    if (h->prm.past > MINCOB_MAX_PAST) return fail(h, MINCOB_E_INVALID, "past > %d not supported", MINCOB_MAX_PAST);
    BatchArgs a = base_args(h);
    a.x = x; a.f_out = f; a.status = status; a.iters = iters; a.evals = evals; a.coeffs = coeffs; a.T = T;
    if ((rc = start_upload(h))) return rc;   // x is already on the device: just get the batch moving
    if (h->upload_in_flight) {          // chunked upload: start as soon as the arrival counter has been reset
        CU(h, cudaStreamWaitEvent(h->stream, h->ev_copy0, 0));
        a.ready = h->ready;
    }
    CU(h, cudaEventRecord(h->ev0, h->stream));
    rc = do_optimize(h, a);
    if (rc) return rc;
    CU(h, cudaEventRecord(h->ev1, h->stream));
    h->timed = true; h->last_launches = 1;
    return 0;
}

int mincob_optimize(mincob_handle h, double *x, double *f, int32_t *status, int32_t *iters, int32_t *evals,
                    double *coeffs, double *T) {
    int rc = have_problems(h, true);
    if (rc) return rc;
    if (!x) return fail(h, MINCOB_E_INVALID, "x must be non-null");
    ON_DEVICE(h);
    const size_t B = h->B, N = h->N, n = N + 3 * (N - 1), S = h->prm.S;
    const size_t nx = B * n * 8, nf = B * 8, ni = B * 4, nc = B * N * 3 * 2 * S * 8, nt = B * N * 8;
    if ((rc = ensure(h, h->b_x, nx)) || (rc = ensure(h, h->b_f, nf)) || (rc = ensure(h, h->b_status, ni)) ||
        (rc = ensure(h, h->b_iters, ni)) || (rc = ensure(h, h->b_evals, ni)) || (rc = ensure(h, h->b_coeffs, nc)) ||
        (rc = ensure(h, h->b_T, nt)))
        return rc;
    if (h->upload_pending) { if ((rc = start_upload(h, x, (double *)h->b_x.p, nx))) return rc; }
    else CU(h, cudaMemcpyAsync(h->b_x.p, x, nx, cudaMemcpyHostToDevice, h->stream));
    rc = mincob_optimize_device(h, (double *)h->b_x.p, (double *)h->b_f.p, (int32_t *)h->b_status.p,
                                (int32_t *)h->b_iters.p, (int32_t *)h->b_evals.p, coeffs ? (double *)h->b_coeffs.p : nullptr,
                                T ? (double *)h->b_T.p : nullptr);
    if (rc) return rc;
    const bool ran = h->timed;  // false when the parameter check short-circuited
    if (ran) CU(h, cudaMemcpyAsync(x, h->b_x.p, nx, cudaMemcpyDeviceToHost, h->stream));
    if (f && ran) CU(h, cudaMemcpyAsync(f, h->b_f.p, nf, cudaMemcpyDeviceToHost, h->stream));
    if (status) CU(h, cudaMemcpyAsync(status, h->b_status.p, ni, cudaMemcpyDeviceToHost, h->stream));
    if (iters && ran) CU(h, cudaMemcpyAsync(iters, h->b_iters.p, ni, cudaMemcpyDeviceToHost, h->stream));
    if (evals && ran) CU(h, cudaMemcpyAsync(evals, h->b_evals.p, ni, cudaMemcpyDeviceToHost, h->stream));
    if (coeffs && ran) CU(h, cudaMemcpyAsync(coeffs, h->b_coeffs.p, nc, cudaMemcpyDeviceToHost, h->stream));
    if (T && ran) CU(h, cudaMemcpyAsync(T, h->b_T.p, nt, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- MINCO building blocks (host pointers) -------------------------------------------------
static int up(mincob_ctx *h, DevBuf &b, const void *src, size_t bytes) {
    int rc = ensure(h, b, bytes);
    if (rc) return rc;
    if (src && bytes) CU(h, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
}
static int down(mincob_ctx *h, void *dst, const DevBuf &b, size_t bytes) {
    if (dst && bytes) CU(h, cudaMemcpyAsync(dst, b.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

// device-pointer forms: enqueue on the handle's stream and return (no copies, no synchronisation)
int mincob_minco_forward_device(mincob_handle h, int B, int N, const double *head, const double *tail, const double *inPs,
                                const double *ts, double *coeffs_asc, double *energy, double *gdC, double *gdT, double *flat) {
    if (!h || !head || !tail || !ts || (N > 1 && !inPs)) return MINCOB_E_INVALID;
    int rc = check_shape(h, B, N, 0);
    if (rc) return rc;
    ON_DEVICE(h);
    MincoArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.N = N;
    a.head = head; a.tail = tail; a.inPs = inPs; a.ts = ts;
    a.coeffs_asc = coeffs_asc; a.energy = energy; a.gdC = gdC; a.gdT = gdT; a.flat = flat;
    return do_minco(h, a, 0);
}

int mincob_minco_propagate_device(mincob_handle h, int B, int N, const double *head, const double *tail, const double *inPs,
                                  const double *ts, const double *gdC, const double *gdT, double *gradByPoints,
                                  double *gradByTimes) {
    if (!h || !head || !tail || !ts || !gdC || !gdT || !gradByTimes || (N > 1 && (!inPs || !gradByPoints)))
        return MINCOB_E_INVALID;
    int rc = check_shape(h, B, N, 0);
    if (rc) return rc;
    ON_DEVICE(h);
    MincoArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.N = N;
    a.head = head; a.tail = tail; a.inPs = inPs; a.ts = ts;
    a.gdC_in = gdC; a.gdT_in = gdT; a.gradByPoints = gradByPoints; a.gradByTimes = gradByTimes;
    return do_minco(h, a, 1);
}

int mincob_minco_forward(mincob_handle h, int B, int N, const double *head, const double *tail, const double *inPs,
                         const double *ts, double *coeffs_asc, double *energy, double *gdC, double *gdT, double *flat) {
    if (!h || !head || !tail || !ts || (N > 1 && !inPs)) return MINCOB_E_INVALID;
    int rc = check_shape(h, B, N, 0);
    if (rc) return rc;
    ON_DEVICE(h);
    const size_t S = h->prm.S, D = 2 * S;
    const size_t nb = (size_t)B * S * 3 * 8, nq = (size_t)B * (N - 1) * 3 * 8, nt = (size_t)B * N * 8,
                 nc = (size_t)B * D * N * 3 * 8, ne = (size_t)B * 8;
    if ((rc = up(h, h->b_m0, head, nb)) || (rc = up(h, h->b_m1, tail, nb)) || (rc = up(h, h->b_m2, inPs, nq)) ||
        (rc = up(h, h->b_m3, ts, nt)) || (rc = ensure(h, h->b_m4, nc)) || (rc = ensure(h, h->b_m5, ne)) ||
        (rc = ensure(h, h->b_m6, nc)) || (rc = ensure(h, h->b_m7, nt)) || (rc = ensure(h, h->b_m8, nc)))
        return rc;
    MincoArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.N = N;
    a.head = (double *)h->b_m0.p; a.tail = (double *)h->b_m1.p; a.inPs = (double *)h->b_m2.p; a.ts = (double *)h->b_m3.p;
    a.coeffs_asc = (double *)h->b_m4.p; a.energy = (double *)h->b_m5.p; a.gdC = (double *)h->b_m6.p;
    a.gdT = (double *)h->b_m7.p; a.flat = (double *)h->b_m8.p;
    if ((rc = do_minco(h, a, 0))) return rc;
    if ((rc = down(h, coeffs_asc, h->b_m4, nc)) || (rc = down(h, energy, h->b_m5, ne)) || (rc = down(h, gdC, h->b_m6, nc)) ||
        (rc = down(h, gdT, h->b_m7, nt)) || (rc = down(h, flat, h->b_m8, nc)))
        return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mincob_minco_propagate(mincob_handle h, int B, int N, const double *head, const double *tail, const double *inPs,
                           const double *ts, const double *gdC, const double *gdT, double *gradByPoints,
                           double *gradByTimes) {
    if (!h || !head || !tail || !ts || !gdC || !gdT || !gradByTimes || (N > 1 && (!inPs || !gradByPoints)))
        return MINCOB_E_INVALID;
    int rc = check_shape(h, B, N, 0);
    if (rc) return rc;
    ON_DEVICE(h);
    const size_t S = h->prm.S, D = 2 * S;
    const size_t nb = (size_t)B * S * 3 * 8, nq = (size_t)B * (N - 1) * 3 * 8, nt = (size_t)B * N * 8,
                 nc = (size_t)B * D * N * 3 * 8;
    if ((rc = up(h, h->b_m0, head, nb)) || (rc = up(h, h->b_m1, tail, nb)) || (rc = up(h, h->b_m2, inPs, nq)) ||
        (rc = up(h, h->b_m3, ts, nt)) || (rc = up(h, h->b_m4, gdC, nc)) || (rc = up(h, h->b_m5, gdT, nt)) ||
        (rc = ensure(h, h->b_m6, nq)) || (rc = ensure(h, h->b_m7, nt)))
        return rc;
    MincoArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.N = N;
    a.head = (double *)h->b_m0.p; a.tail = (double *)h->b_m1.p; a.inPs = (double *)h->b_m2.p; a.ts = (double *)h->b_m3.p;
    a.gdC_in = (double *)h->b_m4.p; a.gdT_in = (double *)h->b_m5.p;
    a.gradByPoints = (double *)h->b_m6.p; a.gradByTimes = (double *)h->b_m7.p;
    if ((rc = do_minco(h, a, 1))) return rc;
    if ((rc = down(h, gradByPoints, h->b_m6, nq)) || (rc = down(h, gradByTimes, h->b_m7, nt))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- feasibility report --------------------------------------------------------------------
int mincob_check_feasibility_device(mincob_handle h, const double *coeffs, const double *T, int samples, double *report) {
    int rc = have_problems(h);
    if (rc) return rc;
    if (!coeffs || !T || !report || samples < 1) return fail(h, MINCOB_E_INVALID, "coeffs, T, report must be non-null and samples >= 1");
    ON_DEVICE(h);
    CheckArgs a;
    memset(&a, 0, sizeof a);
    a.B = h->B; a.N = h->N; a.K = h->K; a.samples = samples;
    a.coeffs = coeffs; a.T = T; a.hpolys = h->hpolys; a.hrows = h->hrows; a.out = report;
    return launched(h, table_for(h->prm.S, a.N)->check(h->stream, h->sm_count, a), "check_kernel");
}

int mincob_check_feasibility(mincob_handle h, const double *coeffs, const double *T, int samples, double *report) {
    int rc = have_problems(h);
    if (rc) return rc;
    if (!coeffs || !T || !report) return fail(h, MINCOB_E_INVALID, "coeffs, T, report must be non-null");
    ON_DEVICE(h);
    const size_t B = h->B, N = h->N, S = h->prm.S;
    const size_t nc = B * N * 3 * 2 * S * 8, nt = B * N * 8, nr = B * 4 * 8;
    if ((rc = up(h, h->b_coeffs, coeffs, nc)) || (rc = up(h, h->b_T, T, nt)) || (rc = ensure(h, h->b_m0, nr))) return rc;
    if ((rc = mincob_check_feasibility_device(h, (const double *)h->b_coeffs.p, (const double *)h->b_T.p, samples,
                                              (double *)h->b_m0.p)))
        return rc;
    if ((rc = down(h, report, h->b_m0, nr))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mincob_max_rates_device(mincob_handle h, const double *coeffs, const double *T, double *rates) {
    int rc = have_problems(h);
    if (rc) return rc;
    if (!coeffs || !T || !rates) return fail(h, MINCOB_E_INVALID, "coeffs, T, rates must be non-null");
    ON_DEVICE(h);
    RateArgs a;
    memset(&a, 0, sizeof a);
    a.B = h->B; a.N = h->N; a.grid = 128;
    a.coeffs = coeffs; a.T = T; a.out = rates;
    return launched(h, table_for(h->prm.S, a.N)->maxrates(h->stream, h->sm_count, a), "maxrate_kernel");
}

int mincob_max_rates(mincob_handle h, const double *coeffs, const double *T, double *rates) {
    int rc = have_problems(h);
    if (rc) return rc;
    if (!coeffs || !T || !rates) return fail(h, MINCOB_E_INVALID, "coeffs, T, rates must be non-null");
    ON_DEVICE(h);
    const size_t nc = (size_t)h->B * h->N * 3 * 2 * h->prm.S * sizeof(double), nt = (size_t)h->B * h->N * sizeof(double);
    const size_t nr = (size_t)h->B * 3 * sizeof(double);
    if ((rc = up(h, h->b_coeffs, coeffs, nc)) || (rc = up(h, h->b_T, T, nt)) || (rc = ensure(h, h->b_m0, nr))) return rc;
    if ((rc = mincob_max_rates_device(h, (const double *)h->b_coeffs.p, (const double *)h->b_T.p, (double *)h->b_m0.p)))
        return rc;
    if ((rc = down(h, rates, h->b_m0, nr))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- measured fp64 ceiling --------------------------------------------------------------------
// Independent DFMA chains on every SM: the fp64-pipe throughput this device sustains, measured with CUDA
// events on the handle's stream.  bench.py divides the optimize kernel's fp64 flop rate by it (the HBM
// roofline BASELINE.json names does not bind this path).
namespace {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *sink, int iters, double a, double b) {
    constexpr int CH = 8;
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = 1.0e-3 * (threadIdx.x + c);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    if (s == 123.456) sink[0] = s;   // never true: keeps the chains alive
}
}  // namespace

int mincob_measure_fp64_peak(mincob_handle h, double *tflops) {
    if (!h || !tflops) return MINCOB_E_INVALID;
    ON_DEVICE(h);
    double *sink = nullptr;
    CU(h, cudaMalloc((void **)&sink, sizeof(double)));
    const int blocks = h->sm_count * 8, threads = 256, iters = 20000;
    cudaEvent_t e0, e1;
    CU(h, cudaEventCreate(&e0));
    CU(h, cudaEventCreate(&e1));
    float best = 0.f;
    for (int rep = 0; rep < 4; ++rep) {          // first pass warms up, best of the next three
        cudaEventRecord(e0, h->stream);
        fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(sink, iters, 0.999999, 1.0e-6);
        cudaEventRecord(e1, h->stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
            return fail(h, MINCOB_E_CUDA, "fp64_peak_kernel: %s", cudaGetErrorString(e));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && (best == 0.f || ms < best)) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    const double flops = 2.0 * 8 * 4 * (double)iters * threads * blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    return 0;
}

// ---- multi-GPU ---------------------------------------------------------------------------
int mincob_nccl_unique_id(void *uid) {
    if (!uid) return MINCOB_E_INVALID;
    if (!nccl_load()) return MINCOB_E_NCCL;
    return g_nccl.get_uid((NcclUid *)uid) == 0 ? 0 : MINCOB_E_NCCL;
}

int mincob_comm_init(mincob_handle h, int nranks, int rank, const void *uid) {
    if (!h || !uid || nranks < 1 || rank < 0 || rank >= nranks) return MINCOB_E_INVALID;
    if (!nccl_load()) return fail(h, MINCOB_E_NCCL, "libnccl not found");
    ON_DEVICE(h);
    NcclUid id;
    memcpy(&id, uid, sizeof id);
    const int rc = g_nccl.init_rank(&h->comm, nranks, id, rank);
    if (rc != 0) return fail(h, MINCOB_E_NCCL, "ncclCommInitRank: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    h->nranks = nranks; h->rank = rank;
    return 0;
}

int mincob_allgather_device(mincob_handle h, const double *send, double *recv, int64_t count) {
    if (!h || !send || !recv || count < 0) return MINCOB_E_INVALID;
    if (!h->comm) return fail(h, MINCOB_E_STATE, "mincob_comm_init has not been called");
    ON_DEVICE(h);
    const int rc = g_nccl.allgather(send, recv, (size_t)count, /*ncclDouble*/ 8, h->comm, h->stream);
    if (rc != 0) return fail(h, MINCOB_E_NCCL, "ncclAllGather: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    return 0;
}

// gather_to_host: coeffs is [nranks*B][N][3][2S] and receives every rank's coefficients; otherwise coeffs is this
// rank's own [B][N][3][2S] and the gathered array stays on the device (mincob_gathered_device).
static int optimize_sharded(mincob_ctx *h, bool gather_to_host, double *x, double *f, int32_t *status, int32_t *iters,
                            int32_t *evals, double *coeffs, double *T) {
    if (!h) return MINCOB_E_INVALID;
    if (!h->comm || h->nranks == 1) return mincob_optimize(h, x, f, status, iters, evals, coeffs, T);
    int rc = have_problems(h, true);
    if (rc) return rc;
    if (!x || (gather_to_host && !coeffs)) return fail(h, MINCOB_E_INVALID, "x and coeffs_all must be non-null");
    ON_DEVICE(h);
    const size_t B = h->B, N = h->N, n = N + 3 * (N - 1), S = h->prm.S;
    const size_t nx = B * n * 8, nf = B * 8, ni = B * 4, cnt = B * N * 3 * 2 * S, nc = cnt * 8, nt = B * N * 8;
    if ((rc = ensure(h, h->b_x, nx)) || (rc = ensure(h, h->b_f, nf)) || (rc = ensure(h, h->b_status, ni)) ||
        (rc = ensure(h, h->b_iters, ni)) || (rc = ensure(h, h->b_evals, ni)) || (rc = ensure(h, h->b_coeffs, nc)) ||
        (rc = ensure(h, h->b_T, nt)) || (rc = ensure(h, h->b_gather, nc * h->nranks)))
        return rc;
    if (h->upload_pending) { if ((rc = start_upload(h, x, (double *)h->b_x.p, nx))) return rc; }
    else CU(h, cudaMemcpyAsync(h->b_x.p, x, nx, cudaMemcpyHostToDevice, h->stream));
    // every rank must reach the collective, also when the parameter check short-circuits
    CU(h, cudaMemsetAsync(h->b_coeffs.p, 0, nc, h->stream));
    rc = mincob_optimize_device(h, (double *)h->b_x.p, (double *)h->b_f.p, (int32_t *)h->b_status.p,
                                (int32_t *)h->b_iters.p, (int32_t *)h->b_evals.p, (double *)h->b_coeffs.p,
                                T ? (double *)h->b_T.p : nullptr);
    if (rc) return rc;
    const bool ran = h->timed;
    if ((rc = mincob_allgather_device(h, (const double *)h->b_coeffs.p, (double *)h->b_gather.p, (int64_t)cnt))) return rc;
    h->gathered_count = (int64_t)cnt * h->nranks;
    if (ran) CU(h, cudaMemcpyAsync(x, h->b_x.p, nx, cudaMemcpyDeviceToHost, h->stream));
    if (f && ran) CU(h, cudaMemcpyAsync(f, h->b_f.p, nf, cudaMemcpyDeviceToHost, h->stream));
    if (status) CU(h, cudaMemcpyAsync(status, h->b_status.p, ni, cudaMemcpyDeviceToHost, h->stream));
    if (iters && ran) CU(h, cudaMemcpyAsync(iters, h->b_iters.p, ni, cudaMemcpyDeviceToHost, h->stream));
    if (evals && ran) CU(h, cudaMemcpyAsync(evals, h->b_evals.p, ni, cudaMemcpyDeviceToHost, h->stream));
    if (gather_to_host) CU(h, cudaMemcpyAsync(coeffs, h->b_gather.p, nc * h->nranks, cudaMemcpyDeviceToHost, h->stream));
    else if (coeffs) CU(h, cudaMemcpyAsync(coeffs, h->b_coeffs.p, nc, cudaMemcpyDeviceToHost, h->stream));
    if (T && ran) CU(h, cudaMemcpyAsync(T, h->b_T.p, nt, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mincob_optimize_sharded(mincob_handle h, double *x, double *f, int32_t *status, int32_t *iters, int32_t *evals,
                            double *coeffs_all, double *T) {
    return optimize_sharded(h, true, x, f, status, iters, evals, coeffs_all, T);
}

int mincob_optimize_sharded_local(mincob_handle h, double *x, double *f, int32_t *status, int32_t *iters, int32_t *evals,
                                  double *coeffs_local, double *T) {
    return optimize_sharded(h, false, x, f, status, iters, evals, coeffs_local, T);
}

int mincob_gathered_device(mincob_handle h, const double **coeffs_all_d, int64_t *count) {
    if (!h || !coeffs_all_d) return MINCOB_E_INVALID;
    if (!h->b_gather.p || h->gathered_count <= 0) return fail(h, MINCOB_E_STATE, "no sharded optimize call has gathered coefficients yet");
    *coeffs_all_d = (const double *)h->b_gather.p;
    if (count) *count = h->gathered_count;
    return 0;
}

#ifdef MINCOB_TIMING
// experiment builds only (tools/): [0] evaluations, [1] time the queue ran dry, [2] kernel end, [3] kernel start (ns),
// [4..67] groups going idle per millisecond after [1]
extern "C" int mincob_debug_counters(mincob_handle h, unsigned long long *out, int n) {
    if (!h || !out || n > 128) return MINCOB_E_INVALID;
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaMemcpy(out, h->total_evals, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}
#endif

int mincob_host_alloc(void **out, uint64_t bytes) {
    if (!out) return MINCOB_E_INVALID;
    *out = nullptr;
    return cudaHostAlloc(out, bytes ? (size_t)bytes : 8, cudaHostAllocDefault) == cudaSuccess ? 0 : MINCOB_E_ALLOC;
}
// page-lock memory the caller already owns (e.g. a POSIX shared-memory segment mapped by every rank of a node, so
// that each rank's device-to-host copy lands directly in the consumer's address space)
int mincob_host_register(void *p, uint64_t bytes) {
    if (!p || !bytes) return MINCOB_E_INVALID;
    return cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable) == cudaSuccess ? 0 : MINCOB_E_CUDA;
}
int mincob_host_unregister(void *p) {
    if (!p) return 0;
    return cudaHostUnregister(p) == cudaSuccess ? 0 : MINCOB_E_CUDA;
}
int mincob_host_free(void *p) {
    if (!p) return 0;
    return cudaFreeHost(p) == cudaSuccess ? 0 : MINCOB_E_CUDA;
}

int mincob_comm_destroy(mincob_handle h) {
    if (!h) return MINCOB_E_INVALID;
    if (h->comm && g_nccl.destroy) g_nccl.destroy(h->comm);
    h->comm = nullptr;
    return 0;
}

}  // extern "C"
