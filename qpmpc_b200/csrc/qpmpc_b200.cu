// qpmpc_b200.cu -- C ABI of the batched MPC engine (see include/qpmpc_b200.h):
// argument checking, kernel-variant dispatch, shared-memory geometry, and the
// host-buffer convenience entry.  No torch, no C++ types across the boundary.

#include "../../include/qpmpc_b200.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mpc_cta_kernel.cuh"
#include "mpc_factor.cuh"
#include "mpc_hessian_tc.cuh"
#include "mpc_host_params.h"
#include "mpc_integrate.cuh"
#include "mpc_launch.cuh"
#include "mpc_lr_kernel.cuh"
#include "mpc_plant.cuh"

using namespace qpmpc;

namespace qpmpc {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1); }
int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}
}  // namespace qpmpc

namespace {


// One CTA per instance (mpc_cta_kernel.cuh).  Shapes whose matrices fit in 227 KB of shared
// memory keep everything there; larger ones keep the vectors in shared memory and the matrices
// in a stream-ordered global-memory workspace (one slice per CTA, a bounded grid striding over
// the batch), so no horizon is refused for its size until the VECTORS outgrow shared memory.
template <typename T>
size_t cta_smem_bytes(int n, int m, int nx) {
    return (size_t)cta_layout(n, m, nx, (int)sizeof(T)).total * sizeof(T);
}
constexpr size_t CTA_SMEM_MAX = 227 * 1024;
constexpr int CTA_MAX_ROWS = 4096;  // the row index field of the pivot key
constexpr int CTA_MAX_VARS = 512;   // largest n accepted (tested up to 256 on the device)

struct CtaPlan {
    size_t smem = 0, ws_bytes = 0;
    int grid = 0;
    void *ws = nullptr;
};
template <typename T>
int plan_cta(const SolveParams &p, cudaStream_t stream, CtaPlan *plan) {
    const CtaLay L = cta_layout(p.n, p.m, p.nx, (int)sizeof(T));
    plan->smem = (size_t)L.total * sizeof(T);
    plan->grid = p.batch;
    const bool forced = env_int("QPMPC_B200_CTA_WORKSPACE", 0) != 0;  // tests: the workspace path on small shapes
    if (plan->smem <= CTA_SMEM_MAX && !forced) return 0;
    plan->smem = (size_t)(L.total - L.oV) * sizeof(T);
    if (plan->smem > CTA_SMEM_MAX || p.m > CTA_MAX_ROWS || p.n > CTA_MAX_VARS) return QPMPC_B200_ESHAPE;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = (int)(CTA_SMEM_MAX / (plan->smem + 1024));
    const int resident = sms * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
    plan->grid = p.batch < resident ? p.batch : resident;
    plan->ws_bytes = (size_t)plan->grid * L.oV * sizeof(T);
    cudaError_t err = cudaMallocAsync(&plan->ws, plan->ws_bytes, stream);
    return err == cudaSuccess ? 0 : (int)err;
}

template <typename T>
int launch_solve_cta(SolveParams p, cudaStream_t stream) {
    p.toeplitz = env_int("QPMPC_B200_NO_TOEPLITZ", 0) == 0;  // the time-invariant condensing path may be used
    CtaPlan plan;
    int rc = plan_cta<T>(p, stream, &plan);
    if (rc) return rc;
    p.workspace = plan.ws;
    auto kern = mpc_solve_cta_kernel<T>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (err == cudaSuccess) {
        int threads = env_int("QPMPC_B200_CTA_THREADS", 256);
        threads = threads < 32 ? 32 : (threads > 256 ? 256 : (threads & ~31));
        kern<<<plan.grid, threads, plan.smem, stream>>>(p);
        count_launch();
        err = cudaGetLastError();
    }
    if (plan.ws) cudaFreeAsync(plan.ws, stream);
    return (int)err;
}

template <typename T>
int launch_condense_cta(SolveParams p, cudaStream_t stream) {
    CtaPlan plan;
    int rc = plan_cta<T>(p, stream, &plan);
    if (rc) return rc;
    p.workspace = plan.ws;
    auto kern = mpc_condense_cta_kernel<T>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (err == cudaSuccess) {
        kern<<<plan.grid, 256, plan.smem, stream>>>(p);
        count_launch();
        err = cudaGetLastError();
    }
    if (plan.ws) cudaFreeAsync(plan.ws, stream);
    return (int)err;
}

// desc.paired is honoured when the rows really split into two halves per step (nc even)
bool rows_paired(const qpmpc_b200_desc *d) {
    return d->paired != 0 && d->nc > 0 && (d->nc & 1) == 0 && env_int("QPMPC_B200_NO_PAIRED", 0) == 0;
}
bool use_cta(int n, int m, Variant *v, bool paired = false) {
    return env_int("QPMPC_B200_FORCE_CTA", 0) != 0 || !pick_variant(n, m, v, paired);
}

template <typename T>
int dispatch_solve(const SolveParams &p, const Variant &v, cudaStream_t s) {
    if (v.paired && v.np == 8) return launch_solve_paired<T, 8>(p, s);
    if (v.paired && v.np == 16) return launch_solve_paired<T, 16>(p, s);
    if (v.paired && v.np == 32) return launch_solve_paired<T, 32>(p, s);
    if (v.np == 8 && v.mr == 2) return launch_solve<T, 8, 2, true>(p, s);
    if (v.np == 8 && v.mr == 4) return launch_solve<T, 8, 4, true>(p, s);
    if (v.np == 16 && v.mr == 2) return launch_solve<T, 16, 2, true>(p, s);
    if (v.np == 16 && v.mr == 4) return launch_solve<T, 16, 4, false>(p, s);
    if (v.np == 32 && v.mr == 2) return launch_solve<T, 32, 2, false>(p, s);
    if (v.np == 32 && v.mr == 4) return launch_solve<T, 32, 4, false>(p, s);
    return QPMPC_B200_ESHAPE;
}

int dispatch_pdip(const SolveParams &p, const Variant &v, int polish, cudaStream_t s) {
    if (v.np == 8 && v.mr == 2) return launch_pdip<double, 8, 2>(p, polish, s);
    if (v.np == 8 && v.mr == 4) return launch_pdip<double, 8, 4>(p, polish, s);
    if (v.np == 16 && v.mr == 2) return launch_pdip<double, 16, 2>(p, polish, s);
    if (v.np == 16 && v.mr == 4) return launch_pdip<double, 16, 4>(p, polish, s);
    if (v.np == 32 && v.mr == 2) return launch_pdip<double, 32, 2>(p, polish, s);
    if (v.np == 32 && v.mr == 4) return launch_pdip<double, 32, 4>(p, polish, s);
    return QPMPC_B200_ESHAPE;
}

template <typename T>
int dispatch_condense(const SolveParams &p, const Variant &v, cudaStream_t s) {
    if (v.np == 8 && v.mr == 2) return launch_condense<T, 8, 2>(p, s);
    if (v.np == 8 && v.mr == 4) return launch_condense<T, 8, 4>(p, s);
    if (v.np == 16 && v.mr == 2) return launch_condense<T, 16, 2>(p, s);
    if (v.np == 16 && v.mr == 4) return launch_condense<T, 16, 4>(p, s);
    if (v.np == 32 && v.mr == 2) return launch_condense<T, 32, 2>(p, s);
    if (v.np == 32 && v.mr == 4) return launch_condense<T, 32, 4>(p, s);
    return QPMPC_B200_ESHAPE;
}

// ---- host-buffer entry: cached device buffers, one set per calling thread ----
// A replayable CUDA graph of one host-buffer solve: the chunked copies and kernels of
// qpmpc_b200_solve_host for one (descriptor, host pointers, chunk) combination.
struct HostGraph {
    qpmpc_b200_desc desc;
    qpmpc_b200_operands in;
    qpmpc_b200_outputs out;
    int chunk = 0;
    int kernels = 0;
    cudaGraphExec_t exec = nullptr;
    unsigned long long last_use = 0;
};

struct HostCache {
    static constexpr int NSTREAM = 3;
    static constexpr int NGRAPH = 8;
    int device = -1;
    cudaStream_t stream[NSTREAM] = {nullptr, nullptr, nullptr};
    cudaEvent_t shared_ready = nullptr, fork = nullptr, join[NSTREAM] = {nullptr, nullptr, nullptr};
    void *buf[OP_COUNT] = {nullptr};
    size_t cap[OP_COUNT] = {0};
    void *U = nullptr, *Z = nullptr;
    int32_t *status = nullptr, *iters = nullptr;
    size_t capU = 0, capZ = 0, capS = 0, capI = 0;
    HostGraph graphs[NGRAPH];
    unsigned long long clock = 0;

    void drop_graphs() {
        for (HostGraph &g : graphs) {
            if (g.exec) cudaGraphExecDestroy(g.exec);
            g = HostGraph();
        }
    }
    // Frees everything (under the device it was created on).
    void release() {
        if (device < 0) return;
        int cur = -1;
        cudaGetDevice(&cur);
        if (cudaSetDevice(device) == cudaSuccess) {
            drop_graphs();
            for (int i = 0; i < NSTREAM; ++i) {
                if (stream[i]) cudaStreamDestroy(stream[i]);
                if (join[i]) cudaEventDestroy(join[i]);
            }
            if (shared_ready) cudaEventDestroy(shared_ready);
            if (fork) cudaEventDestroy(fork);
            for (int o = 0; o < OP_COUNT; ++o)
                if (buf[o]) cudaFree(buf[o]);
            if (U) cudaFree(U);
            if (Z) cudaFree(Z);
            if (status) cudaFree(status);
            if (iters) cudaFree(iters);
        }
        if (cur >= 0) cudaSetDevice(cur);
        *this = HostCache();
    }
    HostCache() = default;
    HostCache(const HostCache &) = default;
    HostCache &operator=(const HostCache &) = default;
    ~HostCache() {
        // thread exit: release unless the CUDA runtime is already being torn down
        if (device >= 0 && cudaFree(nullptr) == cudaSuccess) release();
    }
};
thread_local HostCache g_cache;

// `bytes` fit in the cached buffer; *grew is set when it had to be reallocated.
cudaError_t ensure(void **ptr, size_t *cap, size_t bytes, bool *grew = nullptr) {
    if (bytes <= *cap) return cudaSuccess;
    if (grew) *grew = true;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaSuccess) *cap = bytes;
    return e;
}

// Device address of a page-locked host buffer (nullptr if `p` is pageable or not mapped).
void *mapped_pointer(const void *p) {
    if (!p) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}
bool is_pinned(const void *p) { return !p || mapped_pointer(p) != nullptr; }

// FP64 FMA peak probe: 16 independent dependent chains per thread.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fma(v[i], a, b);
    }
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += v[i];
    if (acc == 123.456) out[0] = acc;  // never true; keeps the chains alive
}

}  // namespace

extern "C" {

int qpmpc_b200_fp64_peak(int device, double *tflops) {
    if (!tflops) return QPMPC_B200_EINVAL;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    double *out = nullptr;
    if ((e = cudaMalloc(&out, 8)) != cudaSuccess) return (int)e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int iters = 4096, grid = sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0);
        dfma_peak_kernel<<<grid, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flops = 2.0 * 16.0 * iters * (double)grid * threads;
        if (rep > 0 && ms > 0.f) best = fmax(best, flops / (ms * 1e-3) / 1e12);
        count_launch();
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(out);
    *tflops = best;
    return (int)cudaGetLastError();
}

int qpmpc_b200_version(void) { return QPMPC_B200_VERSION; }

long long qpmpc_b200_launch_count(void) { return g_launches.load(); }

size_t qpmpc_b200_workspace_bytes(const qpmpc_b200_desc *) { return 0; }

// Capacity of the CTA kernel, assuming nx <= 8: its VECTORS must fit in shared memory (the
// matrices move to a global-memory workspace beyond 227 KB) and a row index in 12 bits.
static size_t cta_vector_bytes(int dtype, int n, int m) {
    const int es = dtype == QPMPC_B200_F64 ? 8 : 4;
    const CtaLay L = cta_layout(n, m, 8, es);
    return (size_t)(L.total - L.oV) * es;
}

int qpmpc_b200_max_vars(int dtype) {
    int n = 32;
    while (n + 1 <= CTA_MAX_VARS && 2 * (n + 1) <= CTA_MAX_ROWS && cta_vector_bytes(dtype, n + 1, 2 * (n + 1)) <= CTA_SMEM_MAX) ++n;  // m = 2 n rows
    return n;
}

int qpmpc_b200_max_rows(int dtype, int n) {
    if (n <= 0 || n > CTA_MAX_VARS || cta_vector_bytes(dtype, n, 0) > CTA_SMEM_MAX) return -1;
    int m = 0;
    while (m + 8 <= CTA_MAX_ROWS && cta_vector_bytes(dtype, n, m + 8) <= CTA_SMEM_MAX) m += 8;
    return m;
}

const char *qpmpc_b200_strerror(int code) {
    switch (code) {
        case 0: return "success";
        case QPMPC_B200_EINVAL: return "invalid argument (null pointer, bad mode or dimension)";
        case QPMPC_B200_ESHAPE: return "N*nu or N*nc exceeds what the kernels hold (see qpmpc_b200_max_vars / _max_rows)";
        case QPMPC_B200_EWEIGHT: return "weights: need w_u > 0 and at least one of w_t, w_x";
        case QPMPC_B200_ENODEVICE: return "no usable CUDA device";
        case QPMPC_B200_EUNSUPPORTED: return "combination not implemented";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

// Does this problem go to the long-horizon kernel (mpc_lr_kernel.cuh)?
static bool takes_lr(const qpmpc_b200_desc *d, const SolveParams &p) {
    const int lr = env_int("QPMPC_B200_LR", -1);
    return d->dtype == QPMPC_B200_F64 && lr != 0 && env_int("QPMPC_B200_FORCE_CTA", 0) == 0 &&
           lr_applicable(p, rows_paired(d)) && (p.n > 16 || (p.n > 8 && env_int("QPMPC_B200_LR16", LR16_DEFAULT) != 0));
}

static int solve_impl(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const qpmpc_b200_outputs *out,
                      const qpmpc_b200_peers *peers, void *stream) {
    int rc = check_desc(d, in);
    if (rc) return rc;
    if (!out) return QPMPC_B200_EINVAL;
    if (!peers && (!out->U || !out->status)) return QPMPC_B200_EINVAL;
    if (peers && (peers->count < 1 || peers->count > 8 || peers->row_offset < 0)) return QPMPC_B200_EINVAL;
    if (d->method != QPMPC_B200_ACTIVE_SET && d->method != QPMPC_B200_PDIP) return QPMPC_B200_EINVAL;
    if (d->batch == 0) return 0;
    SolveParams p;
    fill_params(d, in, &p);
    p.U = out->U;
    p.status = out->status;
    p.iters = out->iters;
    p.Z = out->Z;
    if (peers) {
        p.npeers = peers->count;
        p.row_off = peers->row_offset;
        for (int r = 0; r < peers->count; ++r) {
            if (!peers->U[r]) return QPMPC_B200_EINVAL;
            p.peerU[r] = peers->U[r];
            p.peer_status[r] = peers->status[r];
        }
    }
    Variant v;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (d->method == QPMPC_B200_PDIP) {
        // Interior point (mpc_pdip.cuh): warp kernel only (n <= 32, m <= 128), no fused
        // gather, double precision only -- in single precision the iteration stalls
        // above the engine's fp32 bar on 7-15 % of the humanoid instances (emulator
        // measurement, tests/emu), so it is refused rather than offered.
        if (peers || d->dtype != QPMPC_B200_F64 || !pick_variant(p.n, p.m, &v)) return QPMPC_B200_EUNSUPPORTED;
        p.max_iter = d->max_iter > 0 ? d->max_iter : 50;
        const int polish = (d->flags & QPMPC_B200_FLAG_NO_POLISH) ? 0 : 1;
        return dispatch_pdip(p, v, polish, s);
    }
    // Terminal-cost problems with a long horizon: the structure-exploiting kernel (rank-nx Hessian,
    // Toeplitz G, one CTA of n threads per instance); QPMPC_B200_LR=0 switches it off.
    if (takes_lr(d, p))
        return p.n <= 16   ? launch_solve_lr<double, 16>(p, s)
               : p.n <= 32 ? launch_solve_lr<double, 32>(p, s)
                           : launch_solve_lr<double, 64>(p, s);
    if (use_cta(p.n, p.m, &v, rows_paired(d)))
        return d->dtype == QPMPC_B200_F64 ? launch_solve_cta<double>(p, s) : launch_solve_cta<float>(p, s);
    return d->dtype == QPMPC_B200_F64 ? dispatch_solve<double>(p, v, s) : dispatch_solve<float>(p, v, s);
}

int qpmpc_b200_solve(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const qpmpc_b200_outputs *out,
                     void *stream) {
    return solve_impl(d, in, out, nullptr, stream);
}

int qpmpc_b200_solve_scatter(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const qpmpc_b200_outputs *out,
                             const qpmpc_b200_peers *peers, void *stream) {
    if (!peers) return QPMPC_B200_EINVAL;
    return solve_impl(d, in, out, peers, stream);
}

int qpmpc_b200_condense(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const qpmpc_b200_qp_fields *out,
                        void *stream) {
    int rc = check_desc(d, in);
    if (rc) return rc;
    if (!out) return QPMPC_B200_EINVAL;
    if (d->batch == 0) return 0;
    SolveParams p;
    fill_params(d, in, &p);
    p.P = out->P;
    p.q = out->q;
    p.G = out->G;
    p.h = out->h;
    p.Phi = out->Phi;
    p.Psi = out->Psi;
    p.phi_last = out->phi_last;
    p.psi_last = out->psi_last;
    Variant v;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (use_cta(p.n, p.m, &v)) {
        // Single precision, a horizon that fills a tensor-core tile (32 < n <= 64) and a stage cost:
        // the Psi' Q Psi contraction of P runs on tcgen05 (mpc_hessian_tc.cuh) from the Psi stack
        // the condensing kernel has just written.  QPMPC_B200_HESSIAN_TC=0 keeps the SIMT sum.
        const bool tc = d->dtype == QPMPC_B200_F32 && p.n > 32 && p.n <= 64 && p.has_wx && out->P && out->Psi &&
                        (out->psi_last || !p.has_wt) && env_int("QPMPC_B200_HESSIAN_TC", 1) != 0;
        if (tc) p.P = nullptr, p.skip_P = 1;
        rc = d->dtype == QPMPC_B200_F64 ? launch_condense_cta<double>(p, s) : launch_condense_cta<float>(p, s);
        if (rc || !tc) return rc;
        HessianTcParams hp;
        hp.batch = d->batch, hp.N = d->N, hp.nx = d->nx, hp.n = p.n;
        hp.has_wt = p.has_wt, hp.has_wx = p.has_wx;
        hp.w_t = (float)p.w_t, hp.w_x = (float)p.w_x, hp.w_u = (float)p.w_u;
        hp.Psi = static_cast<const float *>(out->Psi);
        hp.psi_last = static_cast<const float *>(out->psi_last);
        hp.P = static_cast<float *>(out->P);
        const size_t smem = hessian_tc_smem_bytes(d->N, d->nx, p.has_wt, p.has_wx);
        if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
        cudaError_t err = cudaFuncSetAttribute(mpc_hessian_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        mpc_hessian_tc_kernel<<<d->batch, 128, smem, s>>>(hp);
        count_launch();
        return (int)cudaGetLastError();
    }
    return d->dtype == QPMPC_B200_F64 ? dispatch_condense<double>(p, v, s) : dispatch_condense<float>(p, v, s);
}

// ---- shared-model fast path (mpc_factor.cuh) ----------------------------------------------
namespace {
bool shared_mode(int mode, bool optional) {
    return mode == QPMPC_B200_SHARED_LTI || mode == QPMPC_B200_SHARED_LTV || (optional && mode == QPMPC_B200_ABSENT);
}
// NP of the paired variant that serves the factored path, 0 if the problem does not qualify.
int factor_np(const qpmpc_b200_desc *d) {
    if (!d || d->N <= 0 || d->nx <= 0 || d->nu <= 0 || d->nc <= 0 || !rows_paired(d)) return 0;
    if (!shared_mode(d->mode_A, false) || !shared_mode(d->mode_B, false) || !shared_mode(d->mode_C, true) ||
        !shared_mode(d->mode_D, true) || d->method != QPMPC_B200_ACTIVE_SET)
        return 0;
    Variant v;
    if (!pick_variant(d->N * d->nu, d->N * d->nc, &v, true) || !v.paired) return 0;
    return v.np;
}
// scratch behind the record: the condensed fields of the model (P, G, Phi, Psi, phi_N, psi_N)
size_t factor_scratch_elems(const qpmpc_b200_desc *d) {
    const size_t n = (size_t)d->N * d->nu, m = (size_t)d->N * d->nc, nx = d->nx, N = d->N;
    return n * n + m * n + N * nx * nx + N * nx * n + nx * nx + nx * n + 16;
}
bool factor_has_ft(const qpmpc_b200_desc *d) {
    return d->has_wx && d->w_x > 1e-10 && d->mode_targets != QPMPC_B200_VEC_ABSENT &&
           !(d->has_wt && d->w_t > 1e-10 && d->mode_goal == QPMPC_B200_VEC_ABSENT);
}
}  // namespace

size_t qpmpc_b200_factor_bytes(const qpmpc_b200_desc *d) {
    const int np = factor_np(d);
    if (!np) return 0;
    const size_t es = d->dtype == QPMPC_B200_F64 ? 8 : 4;
    return ((size_t)factor_layout(np, d->nx, d->N, factor_has_ft(d)).total + factor_scratch_elems(d)) * es;
}

int qpmpc_b200_factor(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, void *record, void *stream) {
    int rc = check_desc(d, in);
    if (rc) return rc;
    if (!record) return QPMPC_B200_EINVAL;
    const int np = factor_np(d);
    if (!np) return QPMPC_B200_EUNSUPPORTED;
    const size_t es = d->dtype == QPMPC_B200_F64 ? 8 : 4;
    const FactorLay F = factor_layout(np, d->nx, d->N, factor_has_ft(d));
    const size_t n = (size_t)d->N * d->nu, m = (size_t)d->N * d->nc, nx = d->nx, N = d->N;
    // the model's MPCQP fields (one instance), into the scratch behind the record
    char *scr = static_cast<char *>(record) + (size_t)F.total * es;
    qpmpc_b200_qp_fields f;
    f.P = scr, scr += n * n * es;
    f.G = scr, scr += m * n * es;
    f.Phi = scr, scr += N * nx * nx * es;
    f.Psi = scr, scr += N * nx * n * es;
    f.phi_last = scr, scr += nx * nx * es;
    f.psi_last = scr;
    f.q = f.h = nullptr;
    qpmpc_b200_desc d1 = *d;
    d1.batch = 1;
    if ((rc = qpmpc_b200_condense(&d1, in, &f, stream)) != 0) return rc;
    SolveParams sp;
    fill_params(d, in, &sp);
    FactorParams fp;
    fp.N = d->N, fp.nx = d->nx, fp.nu = d->nu, fp.nc = d->nc, fp.n = (int)n, fp.m = (int)m, fp.NP = np;
    fp.has_ft = sp.q_wx, fp.q_wt = sp.q_wt, fp.q_wx = sp.q_wx;
    fp.w_t = sp.w_t, fp.w_x = sp.w_x;
    fp.P = f.P, fp.G = f.G, fp.Phi = f.Phi, fp.Psi = f.Psi, fp.phi_last = f.phi_last, fp.psi_last = f.psi_last;
    fp.C = d->mode_C == QPMPC_B200_ABSENT ? nullptr : in->C;
    fp.stepC = sp.op[OP_C].step;
    fp.record = record;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (d->dtype == QPMPC_B200_F64)
        mpc_factor_kernel<double><<<1, 128, FACTOR_SMEM_BYTES, s>>>(fp);
    else
        mpc_factor_kernel<float><<<1, 128, FACTOR_SMEM_BYTES, s>>>(fp);
    count_launch();
    return (int)cudaGetLastError();
}

static int solve_factored_impl(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const void *record,
                               const qpmpc_b200_outputs *out, const LoopDev *loop, void *stream) {
    int rc = check_desc(d, in);
    if (rc) return rc;
    if (!record || !out || !out->U || !out->status) return QPMPC_B200_EINVAL;
    const int np = factor_np(d);
    if (!np) return QPMPC_B200_EUNSUPPORTED;
    if (d->batch == 0) return 0;
    SolveParams p;
    fill_params(d, in, &p);
    p.U = out->U;
    p.status = out->status;
    p.iters = out->iters;
    p.Z = out->Z;
    // (a long terminal-cost horizon is faster on the structure-exploiting kernel than on the
    // record: measured 62 vs 32 M solves/s at N = 32; QPMPC_B200_FACTORED_LR=0 keeps the record)
    if (!loop && d->method == QPMPC_B200_ACTIVE_SET && takes_lr(d, p) && env_int("QPMPC_B200_FACTORED_LR", 1) != 0)
        return solve_impl(d, in, out, nullptr, stream);
    p.record = record;
    if (loop) p.loop = *loop;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool f64 = d->dtype == QPMPC_B200_F64;
    if (np == 8) return f64 ? launch_solve_pre<double, 8>(p, s) : launch_solve_pre<float, 8>(p, s);
    if (np == 16) return f64 ? launch_solve_pre<double, 16>(p, s) : launch_solve_pre<float, 16>(p, s);
    return f64 ? launch_solve_pre<double, 32>(p, s) : launch_solve_pre<float, 32>(p, s);
}

int qpmpc_b200_solve_factored(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const void *record,
                              const qpmpc_b200_outputs *out, void *stream) {
    return solve_factored_impl(d, in, record, out, nullptr, stream);
}

// The closed loops with a factored model run as ONE launch of the shared-model kernel
// (SolveParams::loop); QPMPC_B200_LOOP_FUSED=0 keeps two launches per cycle.
static bool loop_fused(const void *record, int cycles) {
    return record && cycles > 0 && env_int("QPMPC_B200_LOOP_FUSED", 1) != 0;
}

int qpmpc_b200_integrate(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const void *U, void *X,
                         void *stream) {
    if (!d || !in || !U || !X || !in->A || !in->B || !in->x0) return QPMPC_B200_EINVAL;
    if (d->batch < 0 || d->N <= 0 || d->nx <= 0 || d->nu <= 0) return QPMPC_B200_EINVAL;
    if (d->mode_A == QPMPC_B200_ABSENT || d->mode_B == QPMPC_B200_ABSENT || d->mode_x0 == QPMPC_B200_VEC_ABSENT)
        return QPMPC_B200_EINVAL;
    if (d->batch == 0) return 0;
    IntegrateParams p;
    p.batch = d->batch;
    p.N = d->N;
    p.nx = d->nx;
    p.nu = d->nu;
    p.A = in->A;
    p.B = in->B;
    p.x0 = in->x0;
    p.U = U;
    p.X = X;
    auto ltv = [](int mode) { return mode == QPMPC_B200_SHARED_LTV || mode == QPMPC_B200_BATCH_LTV; };
    auto per = [](int mode) { return mode == QPMPC_B200_BATCH_LTI || mode == QPMPC_B200_BATCH_LTV; };
    p.sA = ltv(d->mode_A) ? d->nx * d->nx : 0;
    p.sB = ltv(d->mode_B) ? d->nx * d->nu : 0;
    p.bA = per(d->mode_A) ? (long long)d->nx * d->nx * (ltv(d->mode_A) ? d->N : 1) : 0;
    p.bB = per(d->mode_B) ? (long long)d->nx * d->nu * (ltv(d->mode_B) ? d->N : 1) : 0;
    p.bx0 = d->mode_x0 == QPMPC_B200_VEC_BATCH ? d->nx : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int threads = 128, grid = (d->batch + threads - 1) / threads;
    if (d->dtype == QPMPC_B200_F64)
        mpc_integrate_kernel<double><<<grid, threads, 0, s>>>(p);
    else
        mpc_integrate_kernel<float><<<grid, threads, 0, s>>>(p);
    count_launch();
    return (int)cudaGetLastError();
}

int qpmpc_b200_pendulum_closed_loop(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in,
                                    const qpmpc_b200_outputs *out, const qpmpc_b200_closed_loop *loop,
                                    void *stream) {
    if (!d || !in || !out || !loop) return QPMPC_B200_EINVAL;
    if (d->nx != 4 || d->nu != 1) return QPMPC_B200_ESHAPE;
    if (d->mode_x0 != QPMPC_B200_VEC_BATCH || d->mode_goal != QPMPC_B200_VEC_BATCH ||
        d->mode_targets != QPMPC_B200_VEC_BATCH)
        return QPMPC_B200_EINVAL;
    if (!in->x0 || !in->goal || !in->targets || !loop->v_target || !out->U || !out->status) return QPMPC_B200_EINVAL;
    if (loop->cycles < 0 || loop->substeps <= 0 || !(loop->dt > 0.0) || !(loop->length > 0.0)) return QPMPC_B200_EINVAL;
    if (d->batch == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PendulumStepParams pp;
    pp.batch = d->batch;
    pp.N = d->N;
    pp.n = d->N * d->nu;
    pp.dt = loop->dt;
    pp.T = loop->sampling_period;
    pp.g = loop->gravity;
    pp.omega2 = loop->gravity / loop->length;
    pp.state = const_cast<void *>(in->x0);
    pp.U = out->U;
    pp.status = out->status;
    pp.v_target = loop->v_target;
    pp.goal = const_cast<void *>(in->goal);
    pp.targets = const_cast<void *>(in->targets);
    pp.unsolved = loop->unsolved;
    pp.upright = loop->upright;
    pp.iters = out->iters;
    pp.iter_sum = nullptr;
    const size_t es = d->dtype == QPMPC_B200_F64 ? 8 : 4;
    const int threads = 128, grid = (d->batch + threads - 1) / threads;
    auto step = [&](int substeps, int slot) {
        pp.substeps = substeps;
        pp.iter_sum = (loop->iterations && out->iters && slot > 0) ? reinterpret_cast<long long *>(loop->iterations) + (slot - 1) : nullptr;
        pp.traj = loop->trajectory ? static_cast<char *>(loop->trajectory) + (size_t)slot * d->batch * 4 * es : nullptr;
        if (d->dtype == QPMPC_B200_F64)
            pendulum_step_kernel<double><<<grid, threads, 0, s>>>(pp);
        else
            pendulum_step_kernel<float><<<grid, threads, 0, s>>>(pp);
        count_launch();
    };
    step(0, 0);  // targets of the first cycle from the initial state
    if (loop_fused(loop->record, loop->cycles)) {
        LoopDev ld;
        std::memset(&ld, 0, sizeof(ld));
        ld.kind = 1, ld.cycles = loop->cycles, ld.substeps = loop->substeps;
        ld.dt = pp.dt, ld.T = pp.T, ld.omega2 = pp.omega2, ld.g = pp.g;
        ld.state = pp.state, ld.v_target = pp.v_target, ld.goal = pp.goal, ld.targets = pp.targets;
        ld.traj = loop->trajectory;
        ld.unsolved = loop->unsolved, ld.upright = loop->upright;
        ld.iter_sum = (loop->iterations && out->iters) ? reinterpret_cast<long long *>(loop->iterations) : nullptr;
        int rc = solve_factored_impl(d, in, loop->record, out, &ld, stream);
        return rc ? rc : (int)cudaGetLastError();
    }
    for (int c = 0; c < loop->cycles; ++c) {
        int rc = loop->record ? qpmpc_b200_solve_factored(d, in, loop->record, out, stream)
                              : qpmpc_b200_solve(d, in, out, stream);
        if (rc) return rc;
        step(loop->substeps, c + 1);
    }
    return (int)cudaGetLastError();
}

int qpmpc_b200_lipm_closed_loop(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in,
                                const qpmpc_b200_outputs *out, const qpmpc_b200_lipm_loop *loop, void *stream) {
    if (!d || !in || !out || !loop) return QPMPC_B200_EINVAL;
    if (d->nx != 3 || d->nu != 1 || d->nc != 2) return QPMPC_B200_ESHAPE;
    if (d->mode_x0 != QPMPC_B200_VEC_BATCH || d->mode_goal != QPMPC_B200_VEC_BATCH || d->mode_e != QPMPC_B200_BATCH_LTV)
        return QPMPC_B200_EINVAL;
    if (!in->x0 || !in->goal || !in->e || !loop->support_foot || !loop->strides || !loop->phase_index ||
        !loop->stride_index || !out->U || !out->status)
        return QPMPC_B200_EINVAL;
    if (loop->cycles < 0 || loop->substeps <= 0 || !(loop->sampling_period > 0.0) || loop->nb_dsp_steps < 0 ||
        loop->nb_ssp_steps <= 0 || 2 * (loop->nb_dsp_steps + loop->nb_ssp_steps) < d->N)
        return QPMPC_B200_EINVAL;  // more than two steps in the receding horizon (:108-111)
    if (d->batch == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    LipmStepParams pp;
    pp.batch = d->batch, pp.N = d->N, pp.n = d->N * d->nu;
    pp.nb_dsp = loop->nb_dsp_steps, pp.nb_ssp = loop->nb_ssp_steps;
    pp.dt = loop->sampling_period / loop->substeps;
    pp.foot_size = loop->foot_size, pp.max_zmp = loop->max_zmp_dist;
    pp.state = const_cast<void *>(in->x0);
    pp.U = out->U, pp.status = out->status;
    pp.support_foot = loop->support_foot, pp.strides = loop->strides;
    pp.phase_index = loop->phase_index, pp.stride_index = loop->stride_index;
    pp.e = const_cast<void *>(in->e), pp.goal = const_cast<void *>(in->goal);
    pp.unsolved = loop->unsolved;
    const size_t es = d->dtype == QPMPC_B200_F64 ? 8 : 4;
    const int threads = 128, grid = (d->batch + threads - 1) / threads;
    auto step = [&](int substeps, int slot) {
        pp.substeps = substeps;
        pp.traj = loop->trajectory ? static_cast<char *>(loop->trajectory) + (size_t)slot * d->batch * 3 * es : nullptr;
        if (d->dtype == QPMPC_B200_F64)
            lipm_step_kernel<double><<<grid, threads, 0, s>>>(pp);
        else
            lipm_step_kernel<float><<<grid, threads, 0, s>>>(pp);
        count_launch();
    };
    step(0, 0);  // bounds and goal of the first cycle
    if (loop_fused(loop->record, loop->cycles)) {
        LoopDev ld;
        std::memset(&ld, 0, sizeof(ld));
        ld.kind = 2, ld.cycles = loop->cycles, ld.substeps = loop->substeps;
        ld.dt = pp.dt, ld.nb_dsp = pp.nb_dsp, ld.nb_ssp = pp.nb_ssp;
        ld.foot_size = pp.foot_size, ld.max_zmp = pp.max_zmp;
        ld.state = pp.state, ld.goal = pp.goal, ld.e = pp.e;
        ld.support_foot = pp.support_foot, ld.strides = pp.strides;
        ld.phase_index = pp.phase_index, ld.stride_index = pp.stride_index;
        ld.traj = loop->trajectory;
        ld.unsolved = loop->unsolved;
        int rc = solve_factored_impl(d, in, loop->record, out, &ld, stream);
        return rc ? rc : (int)cudaGetLastError();
    }
    for (int c = 0; c < loop->cycles; ++c) {
        int rc = loop->record ? qpmpc_b200_solve_factored(d, in, loop->record, out, stream)
                              : qpmpc_b200_solve(d, in, out, stream);
        if (rc) return rc;
        step(loop->substeps, c + 1);
    }
    return (int)cudaGetLastError();
}

int qpmpc_b200_solve_host(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const qpmpc_b200_outputs *out,
                          int device) {
    int rc = check_desc(d, in);
    if (rc) return rc;
    if (!out || !out->U || !out->status) return QPMPC_B200_EINVAL;
    if (d->batch == 0) return 0;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    HostCache &c = g_cache;
    if (c.device != device) {
        c.release();  // buffers, streams and graphs of the device this thread used before
        c.device = device;
        for (int i = 0; i < HostCache::NSTREAM; ++i) {
            if ((e = cudaStreamCreateWithFlags(&c.stream[i], cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
            if ((e = cudaEventCreateWithFlags(&c.join[i], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
        }
        if ((e = cudaEventCreateWithFlags(&c.shared_ready, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
        if ((e = cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
    }
    SolveParams p;
    fill_params(d, in, &p);
    const size_t es = d->dtype == QPMPC_B200_F64 ? 8 : 4;
    const char *host[OP_COUNT] = {(const char *)in->A, (const char *)in->B, (const char *)in->C, (const char *)in->D,
                                  (const char *)in->e, (const char *)in->x0, (const char *)in->goal,
                                  (const char *)in->targets};

    // ZERO-COPY (the default with page-locked buffers): no staging copies at all.  Page-locked
    // host memory is mapped into the device's address space, so the solve kernel's own bulk-TMA
    // staging pulls each CTA's operands over PCIe and its epilogue stores the U rows straight
    // into the caller's buffer: upload, compute and download overlap CTA by CTA inside ONE kernel.
    // Only operands shared by the batch (read by every CTA) are copied to the device first.
    if (env_int("QPMPC_B200_HOST_ZEROCOPY", 1) != 0) {
        void *map_in[OP_COUNT] = {nullptr};
        bool ok = true;
        for (int o = 0; o < OP_COUNT && ok; ++o)
            if (p.op[o].ptr) ok = (map_in[o] = mapped_pointer(host[o])) != nullptr;
        void *mU = ok ? mapped_pointer(out->U) : nullptr, *mS = ok ? mapped_pointer(out->status) : nullptr;
        void *mI = (ok && out->iters) ? mapped_pointer(out->iters) : nullptr;
        void *mZ = (ok && out->Z) ? mapped_pointer(out->Z) : nullptr;
        ok = ok && mU && mS && (!out->iters || mI) && (!out->Z || mZ);
        if (ok) {
            for (int o = 0; o < OP_COUNT; ++o) {
                const OperandView &v = p.op[o];
                if (!v.ptr || v.per_instance) continue;
                // (letting every CTA read small shared operands from host memory instead was measured:
                // config 4 end to end 102 -> 74 M solves/s -- the copy call is cheaper)
                bool moved = false;
                if ((e = ensure(&c.buf[o], &c.cap[o], (size_t)v.sz * es, &moved)) != cudaSuccess) return (int)e;
                if (moved) c.drop_graphs();  // (graphs of the staged path hold the old address)
                e = cudaMemcpyAsync(c.buf[o], host[o], (size_t)v.sz * es, cudaMemcpyHostToDevice, c.stream[0]);
                if (e != cudaSuccess) return (int)e;
                map_in[o] = c.buf[o];
            }
            const qpmpc_b200_operands dev = {map_in[OP_A], map_in[OP_B], map_in[OP_C], map_in[OP_D],
                                             map_in[OP_E], map_in[OP_X0], map_in[OP_GOAL], map_in[OP_TGT]};
            const qpmpc_b200_outputs dout = {mU, (int32_t *)mS, (int32_t *)mI, mZ};
            rc = qpmpc_b200_solve(d, &dev, &dout, c.stream[0]);
            e = cudaStreamSynchronize(c.stream[0]);  // (also after a failed launch: the shared copies are in flight)
            if (rc) return rc;
            return (int)e;
        }
    }
    // Device copies of the operands and outputs, cached per calling thread.
    bool grew = false;
    for (int o = 0; o < OP_COUNT; ++o) {
        const OperandView &v = p.op[o];
        if (!v.ptr) continue;
        const size_t bytes = (size_t)v.sz * (v.per_instance ? d->batch : 1) * es;
        if ((e = ensure(&c.buf[o], &c.cap[o], bytes, &grew)) != cudaSuccess) return (int)e;
    }
    const size_t rU = (size_t)p.n * es, rZ = (size_t)p.m * es;
    if ((e = ensure(&c.U, &c.capU, rU * d->batch, &grew)) != cudaSuccess) return (int)e;
    if ((e = ensure((void **)&c.status, &c.capS, (size_t)d->batch * 4, &grew)) != cudaSuccess) return (int)e;
    if (out->iters && (e = ensure((void **)&c.iters, &c.capI, (size_t)d->batch * 4, &grew)) != cudaSuccess)
        return (int)e;
    const bool wantZ = out->Z && rZ;
    if (wantZ && (e = ensure(&c.Z, &c.capZ, rZ * d->batch, &grew)) != cudaSuccess) return (int)e;
    if (grew) c.drop_graphs();  // they hold the old device addresses

    // Pinned host buffers: the whole pipeline below is captured once into a CUDA graph per
    // (descriptor, pointers) and replayed -- one launch per call instead of ~10 stream
    // operations per chunk, which is what allows chunks small enough to hide the first upload and
    // the last download.  Pageable buffers (no overlap possible anyway) take the direct path.
    bool pinned = env_int("QPMPC_B200_HOST_GRAPH", 1) != 0;
    for (int o = 0; o < OP_COUNT && pinned; ++o)
        if (p.op[o].ptr) pinned = is_pinned(host[o]);
    pinned = pinned && is_pinned(out->U) && is_pinned(out->status) && is_pinned(out->iters) && is_pinned(out->Z);
    int chunk = env_int("QPMPC_B200_HOST_CHUNK", 16384);
    if (chunk < 256) chunk = 256;
    // Chunk schedule: full chunks in the middle, a short ramp at both ends -- nothing overlaps
    // the first upload and the last download, so those chunks are a quarter and a half of a
    // full one (QPMPC_B200_HOST_RAMP=0: uniform chunks).
    int sched[64], nsched = 0;
    {
        const bool ramp = env_int("QPMPC_B200_HOST_RAMP", 1) != 0 && d->batch >= 3 * chunk && chunk >= 1024;
        int left = d->batch;
        const int head[2] = {chunk / 4, chunk / 2};
        int tail = ramp ? chunk / 4 + chunk / 2 : 0;
        if (ramp)
            for (int h : head) sched[nsched++] = h, left -= h;
        while (left - tail > 0 && nsched < 60) {
            const int c0 = left - tail < chunk ? left - tail : chunk;
            sched[nsched++] = c0;
            left -= c0;
        }
        if (ramp) sched[nsched++] = chunk / 2, sched[nsched++] = chunk / 4, left -= tail;
        if (left > 0) sched[nsched - 1] += left;  // (more than 60 chunks: the last one takes the rest)
    }

    // Enqueues the pipeline on the cache's streams: operands shared by the batch go up once and
    // every chunk waits for them; chunks of the batch round-robin over the streams so that the
    // upload of one chunk, the kernel of the previous one and the download of the one before
    // overlap.  Returns the number of kernels enqueued (< 0: an error code of the ABI, with
    // *cuda_err set for CUDA errors).
    auto enqueue = [&](cudaError_t *cuda_err) -> int {
        int kernels = 0;
        bool any_shared = false;
        for (int o = 0; o < OP_COUNT; ++o) {
            const OperandView &v = p.op[o];
            if (!v.ptr || v.per_instance) continue;
            *cuda_err = cudaMemcpyAsync(c.buf[o], host[o], (size_t)v.sz * es, cudaMemcpyHostToDevice, c.stream[0]);
            if (*cuda_err != cudaSuccess) return -1000;
            any_shared = true;
        }
        if (any_shared) cudaEventRecord(c.shared_ready, c.stream[0]);
        int lo = 0;
        for (int idx = 0; idx < nsched; lo += sched[idx], ++idx) {
            const int cnt = sched[idx];
            cudaStream_t s = c.stream[idx % HostCache::NSTREAM];
            if (any_shared && idx % HostCache::NSTREAM != 0 && idx < HostCache::NSTREAM)
                cudaStreamWaitEvent(s, c.shared_ready, 0);
            qpmpc_b200_operands dev;
            const void **devp[OP_COUNT] = {&dev.A, &dev.B, &dev.C, &dev.D, &dev.e, &dev.x0, &dev.goal, &dev.targets};
            for (int o = 0; o < OP_COUNT; ++o) {
                *devp[o] = nullptr;
                const OperandView &v = p.op[o];
                if (!v.ptr) continue;
                if (!v.per_instance) {
                    *devp[o] = c.buf[o];
                    continue;
                }
                const size_t off = (size_t)lo * v.sz * es, bytes = (size_t)cnt * v.sz * es;
                *cuda_err = cudaMemcpyAsync((char *)c.buf[o] + off, host[o] + off, bytes, cudaMemcpyHostToDevice, s);
                if (*cuda_err != cudaSuccess) return -1000;
                *devp[o] = (char *)c.buf[o] + off;
            }
            qpmpc_b200_desc dc = *d;
            dc.batch = cnt;
            qpmpc_b200_outputs dout = {(char *)c.U + rU * lo, c.status + lo, out->iters ? c.iters + lo : nullptr,
                                       wantZ ? (char *)c.Z + rZ * lo : nullptr};
            const int krc = qpmpc_b200_solve(&dc, &dev, &dout, s);
            if (krc) {
                *cuda_err = krc > 0 ? (cudaError_t)krc : cudaSuccess;
                return krc > 0 ? -1000 : krc;
            }
            ++kernels;
            cudaError_t e1 = cudaMemcpyAsync((char *)out->U + rU * lo, dout.U, rU * cnt, cudaMemcpyDeviceToHost, s);
            cudaError_t e2 = cudaMemcpyAsync(out->status + lo, dout.status, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s);
            cudaError_t e3 = cudaSuccess, e4 = cudaSuccess;
            if (dout.iters) e3 = cudaMemcpyAsync(out->iters + lo, dout.iters, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s);
            if (dout.Z) e4 = cudaMemcpyAsync((char *)out->Z + rZ * lo, dout.Z, rZ * cnt, cudaMemcpyDeviceToHost, s);
            for (cudaError_t ei : {e1, e2, e3, e4})
                if (ei != cudaSuccess) {
                    *cuda_err = ei;
                    return -1000;
                }
        }
        return kernels;
    };
    // Waits for everything in flight; the first error wins (never returns with copies or
    // kernels still running on the caller's buffers).
    auto drain = [&](cudaError_t first) -> cudaError_t {
        for (int i = 0; i < HostCache::NSTREAM; ++i) {
            cudaError_t ei = cudaStreamSynchronize(c.stream[i]);
            if (first == cudaSuccess) first = ei;
        }
        return first;
    };

    if (pinned) {
        HostGraph *g = nullptr, *victim = &c.graphs[0];
        for (HostGraph &cand : c.graphs) {
            if (cand.exec && cand.chunk == chunk && !std::memcmp(&cand.desc, d, sizeof(*d)) &&
                !std::memcmp(&cand.in, in, sizeof(*in)) && !std::memcmp(&cand.out, out, sizeof(*out)))
                g = &cand;
            if (cand.last_use < victim->last_use) victim = &cand;
        }
        if (!g) {
            cudaError_t ce = cudaSuccess;
            cudaGraph_t graph = nullptr;
            if ((e = cudaStreamBeginCapture(c.stream[0], cudaStreamCaptureModeThreadLocal)) != cudaSuccess) return (int)e;
            cudaEventRecord(c.fork, c.stream[0]);
            for (int i = 1; i < HostCache::NSTREAM; ++i) cudaStreamWaitEvent(c.stream[i], c.fork, 0);
            const int kernels = enqueue(&ce);
            for (int i = 1; i < HostCache::NSTREAM; ++i) {
                cudaEventRecord(c.join[i], c.stream[i]);
                cudaStreamWaitEvent(c.stream[0], c.join[i], 0);
            }
            e = cudaStreamEndCapture(c.stream[0], &graph);
            if (kernels < 0 || e != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                if (kernels < 0 && kernels != -1000) return kernels;
                return (int)(ce != cudaSuccess ? ce : e);
            }
            if (victim->exec) cudaGraphExecDestroy(victim->exec);
            *victim = HostGraph();
            e = cudaGraphInstantiate(&victim->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return (int)e;
            victim->desc = *d;
            victim->in = *in;
            victim->out = *out;
            victim->chunk = chunk;
            victim->kernels = kernels;
            g = victim;
            g_launches.fetch_sub(kernels);  // counted while capturing; every replay counts them below
        }
        g->last_use = ++c.clock;
        e = cudaGraphLaunch(g->exec, c.stream[0]);
        for (int k = 0; k < g->kernels && e == cudaSuccess; ++k) count_launch();
        cudaError_t es0 = cudaStreamSynchronize(c.stream[0]);
        return (int)(e != cudaSuccess ? e : es0);
    }
    cudaError_t ce = cudaSuccess;
    const int kernels = enqueue(&ce);
    e = drain(ce);
    if (kernels < 0 && kernels != -1000) return kernels;
    return (int)e;
}

}  // extern "C"
