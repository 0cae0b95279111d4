// C-ABI of libpodb200.so: context, memory, POD orchestration, projection / reconstruction.
// See include/podb200.h for the contract and the reference interfaces each entry replaces.
#include "common.cuh"
#include <cfloat>
#include <cmath>
#include <algorithm>
#include <mutex>
#include <functional>
#include <condition_variable>
#include <chrono>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <thread>
#include <atomic>
#include <cstdlib>
#include <vector>

static std::string g_err;
static std::mutex g_err_mu;
const char* pb_set_global_error(const std::string& s) {
    std::lock_guard<std::mutex> l(g_err_mu); g_err = s; return g_err.c_str();
}

int pb_reserve(pb_ctx* ctx, double** p, size_t* cap, size_t elems) {
    if (*cap >= elems && *p) return PB_OK;
    PB_ENTER(ctx);
    if (*p) { PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(*p); *p = nullptr; *cap = 0; }
    cudaError_t e = cudaMalloc(p, elems * sizeof(double));
    if (e != cudaSuccess) { cudaGetLastError(); PB_FAIL(ctx, PB_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", elems * sizeof(double), cudaGetErrorString(e)); }
    *cap = elems;
    return PB_OK;
}
int pb_reserve_int(pb_ctx* ctx, int** p, size_t* cap, size_t elems) {
    if (*cap >= elems && *p) return PB_OK;
    PB_ENTER(ctx);
    if (*p) { PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(*p); *p = nullptr; *cap = 0; }
    cudaError_t e = cudaMalloc(p, elems * sizeof(int));
    if (e != cudaSuccess) { cudaGetLastError(); PB_FAIL(ctx, PB_ERR_NOMEM, "cudaMalloc(%zu bytes) failed", elems * sizeof(int)); }
    *cap = elems;
    return PB_OK;
}

int pbi_smem_optin(pb_ctx* ctx, const void* func, size_t need) {
    static std::mutex mu;
    static std::vector<std::pair<std::pair<const void*, int>, size_t>> done;    // ((function, device), limit)
    std::lock_guard<std::mutex> g(mu);
    size_t limit = 0;
    int device = 0;
    if (ctx) device = ctx->device; else PB_CUDA(ctx, cudaGetDevice(&device));      // (configuration calls pass no context)
    for (auto& e : done) if (e.first.first == func && e.first.second == device) { limit = e.second; break; }
    if (!limit) {
        cudaFuncAttributes a;
        PB_CUDA(ctx, cudaFuncGetAttributes(&a, func));
        int optin = 0;
        PB_CUDA(ctx, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        limit = (size_t)optin - a.sharedSizeBytes;
        PB_CUDA(ctx, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        done.push_back({{func, device}, limit});
    }
    if (need > limit) PB_FAIL(ctx, PB_ERR_SHAPE, "kernel needs %zu bytes of shared memory, the device offers %zu", need, limit);
    return PB_OK;
}

namespace {
// Stage timer: two events on the context stream, NO host synchronisation -- pb_last_ms() waits for the stop event when the
// number is asked for.  (A synchronising timer cost one pipeline bubble per stage: 6 of the 16 host syncs of a C2 step.)
struct Timer {
    pb_ctx* c; int which;
    Timer(pb_ctx* ctx, int w, bool = false) : c(ctx), which(w) {
        c->t_pending[which] = false;
        cudaEventRecord(c->tev[which][0], c->stream);
    }
    void stop() {
        cudaEventRecord(c->tev[which][1], c->stream);
        c->t_pending[which] = true;
    }
};
}  // namespace

extern "C" {

// ---------------------------------------------------------------- lifetime
int pb_create(int device, pb_ctx** out) {
    if (!out) return PB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        pb_set_global_error(std::string("pb_create: no CUDA device (") + cudaGetErrorString(e) + "); libpodb200 has no CPU fallback");
        return PB_ERR_CUDA;
    }
    if (device < 0 || device >= n) { pb_set_global_error("pb_create: bad device index"); return PB_ERR_ARG; }
    pb_ctx* ctx = new pb_ctx();
    ctx->device = device;
    PB_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    PB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        pb_set_global_error("pb_create: libpodb200 is built for sm_100a (Blackwell B200) only");
        delete ctx; return PB_ERR_CUDA;
    }
    PB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    PB_CUDA(ctx, cudaEventCreate(&ctx->ev0)); PB_CUDA(ctx, cudaEventCreate(&ctx->ev1));
    PB_CUDA(ctx, cudaEventCreate(&ctx->ev2)); PB_CUDA(ctx, cudaEventCreate(&ctx->ev3));
    PB_CUDA(ctx, cudaEventCreate(&ctx->evk0)); PB_CUDA(ctx, cudaEventCreate(&ctx->evk1));
    PB_CUDA(ctx, cudaEventCreate(&ctx->evr0)); PB_CUDA(ctx, cudaEventCreate(&ctx->evr1));
    for (int w = 0; w < PB_T_COUNT; ++w) { PB_CUDA(ctx, cudaEventCreate(&ctx->tev[w][0])); PB_CUDA(ctx, cudaEventCreate(&ctx->tev[w][1])); }
    PB_CUDA(ctx, cudaMalloc(&ctx->d_int, 256 * sizeof(int)));
    PB_CUDA(ctx, cudaMemset(ctx->d_int, 0, 256 * sizeof(int)));
    PB_CUDA(ctx, cudaMallocHost(&ctx->h_int, 256 * sizeof(int)));
    *out = ctx;
    return PB_OK;
}

int pb_destroy(pb_ctx* ctx) {
    if (!ctx) return PB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    void* bufs[] = {ctx->ws, ctx->flush, ctx->At, ctx->At2, ctx->sig, ctx->nrm, ctx->sig2, ctx->nrm2, ctx->perm,
                    ctx->perm2, ctx->Bmat, ctx->Wk, ctx->small, ctx->qr_W, ctx->qr_scr, ctx->qr_scr2, ctx->qr_stack, ctx->Gbuf, ctx->G2buf, ctx->Ybuf, ctx->ATbuf, ctx->Uf, ctx->d_int, ctx->fp_ibuf,
                    ctx->fp_buf[0], ctx->fp_buf[1], ctx->fp_buf[2], ctx->fp_buf[3], ctx->fp_buf[4], ctx->fp_buf[5], ctx->fp_buf[6], ctx->fp_buf[7]};
    for (void* b : bufs) if (b) cudaFree(b);
    if (ctx->h_int) cudaFreeHost(ctx->h_int);
    if (ctx->h_sig) cudaFreeHost(ctx->h_sig);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->ev2); cudaEventDestroy(ctx->ev3);
    cudaEventDestroy(ctx->evk0); cudaEventDestroy(ctx->evk1); cudaEventDestroy(ctx->evr0); cudaEventDestroy(ctx->evr1);
    for (int w = 0; w < PB_T_COUNT; ++w) { cudaEventDestroy(ctx->tev[w][0]); cudaEventDestroy(ctx->tev[w][1]); }
    if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); for (auto& e : ctx->chunk_ev) if (e) cudaEventDestroy(e); }
    if (ctx->Gpart) cudaFree(ctx->Gpart);
    for (int s = 0; s < PB_STAGE_SLOTS; ++s) { if (ctx->stage[s]) cudaFreeHost(ctx->stage[s]); if (ctx->stage_ev[s]) cudaEventDestroy(ctx->stage_ev[s]); }
    if (ctx->slab) cudaFree(ctx->slab);
    if (ctx->Rbuf) cudaFree(ctx->Rbuf);
    if (ctx->Pbuf) cudaFree(ctx->Pbuf);
    pbi_train_free(ctx);
    pbi_gram_plan_free(ctx);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return PB_OK;
}

const char* pb_last_error(const pb_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> l(g_err_mu); return g_err.c_str();
}
int pb_set_stream(pb_ctx* ctx, void* s) { PB_ENTER(ctx); ctx->stream = s ? (cudaStream_t)s : ctx->own_stream; return PB_OK; }
int pb_sync(pb_ctx* ctx) {
    PB_ENTER(ctx);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->fork_pending) { ctx->fork_pending = false; PB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream)); }
    return PB_OK;
}
int pb_device_info(pb_ctx* ctx, int* sm, int* maj, int* min_, size_t* mem) {
    PB_ENTER(ctx);
    cudaDeviceProp p; PB_CUDA(ctx, cudaGetDeviceProperties(&p, ctx->device));
    if (sm) *sm = p.multiProcessorCount; if (maj) *maj = p.major; if (min_) *min_ = p.minor; if (mem) *mem = p.totalGlobalMem;
    return PB_OK;
}
int pb_last_ms(pb_ctx* ctx, int which, double* ms) {
    PB_ENTER(ctx); if (which < 0 || which >= PB_T_COUNT || !ms) return PB_ERR_ARG;
    if (which == PB_T_TN_KERNEL && ctx->tn_timing_pending) {
        PB_CUDA(ctx, cudaEventSynchronize(ctx->evk1));
        float t = 0; cudaEventElapsedTime(&t, ctx->evk0, ctx->evk1); ctx->ms[PB_T_TN_KERNEL] = t;
        ctx->tn_timing_pending = false;
    }
    if (ctx->t_pending[which]) {
        PB_CUDA(ctx, cudaEventSynchronize(ctx->tev[which][1]));
        float t = 0; PB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->tev[which][0], ctx->tev[which][1]));
        ctx->ms[which] = t; ctx->t_pending[which] = false;
    }
    *ms = ctx->ms[which]; return PB_OK;
}
int pb_timer_start(pb_ctx* ctx) { PB_ENTER(ctx); PB_CUDA(ctx, cudaEventRecord(ctx->evr0, ctx->stream)); return PB_OK; }
int pb_timer_stop(pb_ctx* ctx, double* ms) {
    PB_ENTER(ctx); if (!ms) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaEventRecord(ctx->evr1, ctx->stream));
    PB_CUDA(ctx, cudaEventSynchronize(ctx->evr1));
    float t = 0; PB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->evr0, ctx->evr1)); *ms = t; return PB_OK;
}
int64_t pb_kernel_launches(pb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------- memory
int pb_malloc(pb_ctx* ctx, size_t bytes, void** dev) {
    PB_ENTER(ctx); if (!dev) return PB_ERR_ARG;
    cudaError_t e = cudaMalloc(dev, bytes ? bytes : 8);
    if (e != cudaSuccess) { cudaGetLastError(); PB_FAIL(ctx, PB_ERR_NOMEM, "pb_malloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    return PB_OK;
}
int pb_free(pb_ctx* ctx, void* dev) { PB_ENTER(ctx); if (dev) { cudaStreamSynchronize(ctx->stream); PB_CUDA(ctx, cudaFree(dev)); } return PB_OK; }
int pb_memset(pb_ctx* ctx, void* dev, int v, size_t bytes) { PB_ENTER(ctx); PB_CUDA(ctx, cudaMemsetAsync(dev, v, bytes, ctx->stream)); return PB_OK; }
int pb_upload_async(pb_ctx* ctx, void* dev, const void* host, size_t bytes) {
    PB_ENTER(ctx); PB_CUDA(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream)); return PB_OK;
}
int pb_download_async(pb_ctx* ctx, void* host, const void* dev, size_t bytes) {
    PB_ENTER(ctx); PB_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream)); return PB_OK;
}
// D2H of a finished result on the context's COPY stream: ordered after the work queued so far, overlapping what is
// queued next (e.g. V on its way to the host while the projection V^T U runs).  pb_sync waits for it as well.  The
// source buffer must not be rewritten before pb_sync.
int pb_download_forked(pb_ctx* ctx, void* host, const void* dev, size_t bytes) {
    PB_ENTER(ctx); if (!host || !dev) return PB_ERR_ARG;
    if (!ctx->copy_stream) {
        PB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int c = 0; c < 16; ++c) PB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->chunk_ev[c], cudaEventDisableTiming));
    }
    PB_CUDA(ctx, cudaEventRecord(ctx->chunk_ev[14], ctx->stream));
    PB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[14], 0));
    PB_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    ctx->fork_pending = true;
    return PB_OK;
}
int pb_upload(pb_ctx* ctx, void* dev, const void* host, size_t bytes) { PB_TRY(pb_upload_async(ctx, dev, host, bytes)); return pb_sync(ctx); }
namespace { int download_staged(pb_ctx* ctx, char* host, const char* dev, size_t bytes); bool host_is_pinned(const void* p); }
// D2H + sync.  Large pageable destinations (an ordinary numpy array) go through the pinned staging ring: the copy engine
// fills slot s + 1 while host threads move slot s into the destination (the driver's own pageable path is single-threaded).
int pb_download(pb_ctx* ctx, void* host, const void* dev, size_t bytes) {
    if (ctx && host && dev && bytes >= ((size_t)8 << 20) && !host_is_pinned(host)) {
        PB_ENTER(ctx);
        return download_staged(ctx, static_cast<char*>(host), static_cast<const char*>(dev), bytes);
    }
    PB_TRY(pb_download_async(ctx, host, dev, bytes)); return pb_sync(ctx);
}
int pb_slab(pb_ctx* ctx, size_t bytes, void** dev) {
    PB_ENTER(ctx); if (!dev) return PB_ERR_ARG;
    size_t elems = (bytes + 7) / 8; if (elems == 0) elems = 1;
    PB_TRY(pb_reserve(ctx, &ctx->slab, &ctx->slab_cap, elems));
    *dev = ctx->slab;
    return PB_OK;
}
int pb_host_alloc(size_t bytes, void** host) {
    if (!host) return PB_ERR_ARG;
    cudaError_t e = cudaMallocHost(host, bytes ? bytes : 8);
    if (e != cudaSuccess) { cudaGetLastError(); pb_set_global_error(std::string("pb_host_alloc: ") + cudaGetErrorString(e)); return PB_ERR_NOMEM; }
    return PB_OK;
}
int pb_host_free(void* host) { if (host) cudaFreeHost(host); return PB_OK; }

// ---------------------------------------------------------------- products
int pb_gram(pb_ctx* ctx, const double* A, int64_t rows, int64_t cols, int64_t lda, double* G) {
    PB_ENTER(ctx);
    if (lda < cols) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_gram: lda %lld < n_cols %lld", (long long)lda, (long long)cols);
    Timer t(ctx, PB_T_GRAM);
    ctx->gram_rows = rows; ctx->refine_rows_fixed = 0;     // a new Gram: the override (pb_pod_set_refine_rows) must be renewed
    PB_TRY(pbi_gemm_tn(ctx, A, lda, cols, A, lda, cols, rows, G, cols, 3));   // sym | time the main kernel
    t.stop();
    return PB_OK;
}
int pb_gemm_tn(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka, const double* B, int64_t ldb, int64_t kb,
               int64_t rows, double* C) {
    PB_ENTER(ctx);
    return pbi_gemm_tn(ctx, A, lda, ka, B, ldb, kb, rows, C, kb, 0);
}
int pb_gemm_nn(pb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t m, const double* B, int64_t ldb,
               int64_t k, double* Y, int64_t ldy) {
    PB_ENTER(ctx);
    return pbi_gemm_nn(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, nullptr, 0, nullptr);
}

// ---------------------------------------------------------------- POD
static int finish_rank(pb_ctx* ctx, const double* sig_dev, int64_t count, int avail, double eps, int n_L_in, int* n_L) {
    int nl = 0;
    if (n_L_in > 0) {
        if (n_L_in > avail) PB_FAIL(ctx, PB_ERR_SHAPE, "requested n_L=%d exceeds the numerical rank %d", n_L_in, avail);
        nl = n_L_in;
    } else {
        PB_TRY(pbi_rank(ctx, sig_dev, count, eps, &nl));     // synchronises the stream: a pending sigma copy has landed too
        if (nl > avail) nl = avail;
        if (ctx->sigma_pending) {
            ctx->sigma_host.assign((size_t)ctx->sig_total, 0.);
            memcpy(ctx->sigma_host.data(), ctx->h_sig, sizeof(double) * (size_t)std::min(ctx->sig_count, ctx->sig_total));
            ctx->sigma_pending = false;
        }
    }
    ctx->n_L = nl; *n_L = nl;
    return PB_OK;
}

static int fetch_sigma(pb_ctx* ctx, const double* sig_dev, int64_t count, int64_t total) {
    // asynchronous copy into pinned memory: a copy into the pageable vector would block the host until the stream drains.
    // Copies are ordered on the stream, so a pending earlier one needs no wait -- unless the buffer is about to be replaced
    if (ctx->h_sig_cap < (size_t)total) {
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->h_sig) cudaFreeHost(ctx->h_sig);
        ctx->h_sig = nullptr; ctx->h_sig_cap = 0;
        PB_CUDA(ctx, cudaMallocHost(&ctx->h_sig, sizeof(double) * (size_t)(total + 64)));
        ctx->h_sig_cap = (size_t)total + 64;
    }
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->h_sig, sig_dev, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sig_count = count; ctx->sig_total = total; ctx->sigma_pending = true;
    return PB_OK;
}

int pb_pod_from_gram(pb_ctx* ctx, const double* G, int64_t m, double eps, int n_L_in, int mode, int* n_L, int* n_keep) {
    PB_ENTER(ctx); if (!n_L) return PB_ERR_ARG;
    if (mode == PB_POD_TSQR) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod_from_gram: the TSQR mode starts from pb_tsqr_r / pb_pod_from_r, not from a Gram matrix");
    if (mode < PB_POD_GRAM || mode > PB_POD_ACCURATE) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod: unknown mode %d", mode);
    if (!(eps >= 0.) || eps >= 1.) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod: eps must be in [0, 1)");
    Timer t(ctx, PB_T_EIG);
    ctx->m = m; ctx->mode = mode; ctx->pass2_done = false;
    ctx->refine_steps = 0; ctx->refine_est = 0.; ctx->ms[PB_T_REFINE] = 0.; ctx->t_pending[PB_T_REFINE] = false;
    PB_TRY(pb_reserve(ctx, &ctx->sig, &ctx->sig_cap, (size_t)m + 32));
    PB_TRY(pb_reserve(ctx, &ctx->nrm, &ctx->nrm_cap, (size_t)m + 32));
    PB_TRY(pb_reserve_int(ctx, &ctx->perm, &ctx->perm_cap, (size_t)m + 64));
    // eps == 0 asks for every mode (pod.py default): keep the full factor as well
    const bool full = (mode == PB_POD_ACCURATE) || (eps == 0. && n_L_in == 0);
    const double tol_rel = full ? 0. : 4. * DBL_EPSILON;
    const double shift_rel = full ? 2. * (double)m * DBL_EPSILON : 0.;
    {   // measured on B200 (profiles/r01_notes.md): loosening the first-pass rotation threshold of the accurate
        // mode to 1e-9 / 1e-6 costs 4 / 6 digits of subspace angle and saves < 15 % -- off unless asked for
        static const double tol1 = []{ const char* e = getenv("PB_ACC_TOL1"); return e ? atof(e) : 0.; }();
        if (mode == PB_POD_ACCURATE && tol1 > 0.) ctx->eig_jacobi_tol = tol1;
    }
    {   // a second pass re-orthogonalises: when the first factor is a single 32-vector pair, its in-kernel solve
        // (iterated to the rotation threshold) is taken as is, without the confirming sweep
        ctx->eig_trust_single_pair = (mode != PB_POD_GRAM);
    }
    PB_TRY(pbi_eig_psd(ctx, G, m, tol_rel, shift_rel, &ctx->At, &ctx->At_cap, ctx->sig, ctx->nrm, ctx->perm,
                       &ctx->r, &ctx->sweeps1, 40));
    ctx->n_keep = (mode == PB_POD_GRAM) ? 0 : ctx->r;
    if (n_keep) *n_keep = ctx->n_keep;
    if (ctx->n_keep > 0) {
        // a second pass follows and decides sigma and n_L from the rotated block: no rank kernel, no host round trip
        // here.  *n_L is provisional (the numerical rank) until pb_pod_pass2_finish.
        ctx->n_L = ctx->r; *n_L = ctx->r;
    } else {
        PB_TRY(fetch_sigma(ctx, ctx->sig, m, m));
        PB_TRY(finish_rank(ctx, ctx->sig, m, ctx->r, eps, n_L_in, n_L));
    }
    t.stop();
    return PB_OK;
}

// ---- subspace refinement of the gram2 mode ---------------------------------------------------------------------
// The first Gram pass resolves the right singular vectors z_k only to eps lambda_1 / lambda_k: z_k leaks into the
// directions the numerical-rank block W_r does not span (singular values below sqrt(4 eps) sigma_1, the pivoted
// Cholesky's stopping threshold), and a mode v_k = U z_k / sigma_k inherits that leak damped by sigma_tail / sigma_k.
// Those components lie OUTSIDE span(U W_r), so the second Gram pass cannot remove them (north_star's tolerance: 1e-8).
// A-priori estimate, x = sigma_1 / sigma_nL (set by eps: ~1e5 at eps = 1e-10), m = order of the Gram, n = rows the
// Gram was accumulated over on this GPU (its rounding error grows ~ n^1.2: 9.3e-10 -> 2.9e-8 from 65 536 to 1 M rows):
//   tall: 0.54 eps x^3 sqrt(4 eps) max(1, n/65536)^1.2 max(1, m/1024)^0.8      wide: 0.4 eps x^2  (no damping)
// fitted as an upper envelope (x 3 at the anchor) of the unrefined angles measured on B200 (profiles/r02_gram2_angles.txt):
// DCT decay 1/4 at 65 536 x 1024 / 1 M x 1024 / 200 000 x 2048 / 1.25 M x 4096: 9.3e-10 / 2.9e-8 / 6.5e-9 / 4.6e-8
// (estimates 2.8e-9 / 7.8e-8 / 1.9e-8 / 2.9e-7), C2 2.3e-10 (1.8e-9), Shekel 400 x 300 1.2e-9 (3e-9), Shekel C1 wide
// 4.7e-7 (2e-6), Burgers 256 x 1200 wide 1.2e-8 (3e-8).  When it exceeds
// 1e-8 ONE step of subspace iteration with the matrix itself follows: V~ = A (A^T V) Sigma^-2, V = V~ R^-1 with
// R^T R = V~^T V~ (V~ is nearly orthonormal, so the Cholesky QR is safe); the out-of-span error shrinks by
// (sigma_tail/sigma_nL)^2 <= 1e-5.  PB_GRAM2_REFINE=0/1 forces it off / on.
static int refine_steps_needed(pb_ctx* ctx) {
    static const int env = []{ const char* e = getenv("PB_GRAM2_REFINE"); return e ? atoi(e) : -1; }();
    if (ctx->mode != PB_POD_GRAM2 || ctx->n_L <= 0) return 0;
    if (env == 0) return 0;
    if (env == 1) return 1;
    if (ctx->sigma_host.size() < (size_t)ctx->n_L) return 0;
    const double s1 = ctx->sigma_host[0], sl = ctx->sigma_host[ctx->n_L - 1];
    if (!(sl > 0.) || !(s1 > 0.)) return 0;
    const double x = s1 / sl;
    const int64_t est_rows = ctx->refine_rows_fixed > 0 ? ctx->refine_rows_fixed : ctx->gram_rows;
    const double size_f = pow(std::max(1., (double)est_rows / 65536.), 1.2) * pow(std::max(1., (double)ctx->m / 1024.), 0.8);
    const double est = ctx->wide ? 0.4 * DBL_EPSILON * x * x
                                 : 0.54 * DBL_EPSILON * x * x * x * sqrt(4. * DBL_EPSILON) * size_f;
    ctx->refine_est = est;
    return est > 1e-8 ? 1 : 0;
}

// in-place lower Cholesky of a small SPD matrix (host) and the inverse of its transpose factor: on return
// Rinv (n x n, row-major, upper triangular) satisfies (G = R^T R)  V~ Rinv orthonormal
static bool host_chol_rinv(std::vector<double>& G, int n, std::vector<double>& Rinv) {
    // R upper triangular with G = R^T R (row-oriented Cholesky)
    std::vector<double> R((size_t)n * n, 0.);
    for (int i = 0; i < n; ++i) {
        for (int j = i; j < n; ++j) {
            double s = G[(size_t)i * n + j];
            for (int k = 0; k < i; ++k) s -= R[(size_t)k * n + i] * R[(size_t)k * n + j];
            if (j == i) { if (!(s > 0.)) return false; R[(size_t)i * n + i] = sqrt(s); }
            else R[(size_t)i * n + j] = s / R[(size_t)i * n + i];
        }
    }
    Rinv.assign((size_t)n * n, 0.);
    for (int j = 0; j < n; ++j) {                    // column j of R^-1 by back substitution
        Rinv[(size_t)j * n + j] = 1. / R[(size_t)j * n + j];
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.;
            for (int k = i + 1; k <= j; ++k) s += R[(size_t)i * n + k] * Rinv[(size_t)k * n + j];
            Rinv[(size_t)i * n + j] = -s / R[(size_t)i * n + i];
        }
    }
    return true;
}

namespace {
// B[i][j] *= 1 / sig[j]^2
__global__ void scale_cols_inv_sq_kernel(double* __restrict__ B, int64_t n, int k, const double* __restrict__ sig) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * k) return;
    const double s = sig[idx % k];
    B[idx] = s > 0. ? B[idx] / (s * s) : 0.;
}
}  // namespace

int pb_pod_set_refine_rows(pb_ctx* ctx, int64_t rows) {
    PB_ENTER(ctx); if (rows < 0) return PB_ERR_ARG;
    ctx->refine_rows_fixed = rows;
    return PB_OK;
}
int pb_pod_refine_steps(pb_ctx* ctx, int* steps) {
    PB_ENTER(ctx); if (!steps) return PB_ERR_ARG;
    *steps = ctx->pass2_done ? ctx->refine_steps : 0;
    return PB_OK;
}

int pb_pod_refine_estimate(pb_ctx* ctx, double* est) {
    PB_ENTER(ctx); if (!est) return PB_ERR_ARG;
    *est = ctx->pass2_done ? ctx->refine_est : 0.;
    return PB_OK;
}

// Tall case, row-sharded form.  B (m x n_L, ld n_L) = A^T V summed over all rows (pb_project + all-reduce).  Forms
// V~ = A_local B Sigma^-2 in the context (rows x n_L) and G3 (n_L x n_L) = V~^T V~ over the local rows.
int pb_pod_refine_apply(pb_ctx* ctx, const double* A, int64_t rows, int64_t lda, double* B, double* G3) {
    PB_ENTER(ctx); if (!A || !B || !G3) return PB_ERR_ARG;
    const int nl = ctx->n_L; const int64_t m = ctx->m;
    if (!ctx->pass2_done || nl <= 0) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod_refine_apply: no second-pass decomposition");
    const double* sig = ctx->sig2;
    scale_cols_inv_sq_kernel<<<(unsigned)((m * nl + 255) / 256), 256, 0, ctx->stream>>>(B, m, nl, sig);
    PB_LAUNCH_CHECK(ctx);
    PB_TRY(pb_reserve(ctx, &ctx->Rbuf, &ctx->R_cap, (size_t)rows * nl));
    PB_TRY(pbi_gemm_nn(ctx, A, lda, rows, m, B, nl, nl, ctx->Rbuf, nl, nullptr, 0, nullptr));
    PB_TRY(pbi_gemm_tn(ctx, ctx->Rbuf, nl, nl, ctx->Rbuf, nl, nl, rows, G3, nl, 1));
    return PB_OK;
}

// G3 = the all-reduced V~^T V~.  V (rows x n_L, ldv) = V~ R^-1.
int pb_pod_refine_finish(pb_ctx* ctx, const double* G3, int64_t rows, double* V, int64_t ldv) {
    PB_ENTER(ctx); if (!G3 || !V) return PB_ERR_ARG;
    const int nl = ctx->n_L;
    if (!ctx->Rbuf || nl <= 0) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod_refine_finish: call pb_pod_refine_apply first");
    std::vector<double> g((size_t)nl * nl), rinv;
    PB_CUDA(ctx, cudaMemcpyAsync(g.data(), G3, sizeof(double) * g.size(), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!host_chol_rinv(g, nl, rinv)) PB_FAIL(ctx, PB_ERR_NOT_CONVERGED, "pb_pod_refine_finish: refined block is rank deficient");
    PB_TRY(pb_reserve(ctx, &ctx->Bmat, &ctx->B_cap, (size_t)nl * nl));
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->Bmat, rinv.data(), sizeof(double) * rinv.size(), cudaMemcpyHostToDevice, ctx->stream));
    PB_TRY(pbi_gemm_nn(ctx, ctx->Rbuf, nl, rows, nl, ctx->Bmat, nl, nl, V, ldv, nullptr, 0, nullptr));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // rinv (host) is read by the async copy above
    return PB_OK;
}

int pb_pod_pass2_gram(pb_ctx* ctx, const double* A, int64_t rows, int64_t lda, double* Y, double* G2) {
    PB_ENTER(ctx);
    const int keep = ctx->n_keep; const int64_t m = ctx->m;
    if (keep <= 0) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod_pass2_gram: the last decomposition has no second pass (mode GRAM)");
    Timer t2(ctx, PB_T_PASS2);       // stopped by pb_pod_pass2_finish (a sharded caller sums G2 over the GPUs in between)
    ctx->pass2_t_started = true;
    PB_TRY(pb_reserve(ctx, &ctx->Wk, &ctx->Wk_cap, (size_t)m * keep));
    PB_TRY(pbi_factor_to_right(ctx, ctx->At, m, keep, ctx->perm, ctx->nrm, nullptr, ctx->Wk, keep));
    PB_TRY(pbi_gemm_nn(ctx, A, lda, rows, m, ctx->Wk, keep, keep, Y, keep, nullptr, 0, nullptr));
    PB_TRY(pbi_gemm_tn(ctx, Y, keep, keep, Y, keep, keep, rows, G2, keep, 1));
    return PB_OK;
}

int pb_pod_pass2_finish(pb_ctx* ctx, const double* G2, double eps, int n_L_in, int* n_L) {
    PB_ENTER(ctx); if (!n_L) return PB_ERR_ARG;
    const int keep = ctx->n_keep;
    if (keep <= 0) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod_pass2_finish: no second pass pending");
    PB_TRY(pb_reserve(ctx, &ctx->sig2, &ctx->sig2_cap, (size_t)keep + 32));
    PB_TRY(pb_reserve(ctx, &ctx->nrm2, &ctx->nrm2_cap, (size_t)keep + 32));
    PB_TRY(pb_reserve_int(ctx, &ctx->perm2, &ctx->perm2_cap, (size_t)keep + 64));
    // G2 = D (I + E) D is graded and nearly diagonal: no shift, no truncation (relative accuracy)
    bool done = false;
    static const bool queued_ok = []{ const char* e = getenv("PB_EIG_SMALL_QUEUED"); return !e || atoi(e) != 0; }();
    if (queued_ok && keep <= 32 && n_L_in == 0) {
        // one host round trip for the whole small decomposition (Cholesky, single-pair Jacobi, norms, sort, rank rule)
        int rq = pbi_eig_psd_small_queued(ctx, G2, keep, &ctx->At2, &ctx->At2_cap, ctx->sig2, ctx->nrm2, ctx->perm2, eps);
        if (rq < 0) return rq;
        if (rq == PB_OK) {
            PB_TRY(fetch_sigma(ctx, ctx->sig2, keep, ctx->m));
            PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            const int r = ctx->h_int[0], nl = ctx->h_int[2];
            if (ctx->h_int[5] == 1 && ctx->h_int[4] == 1 && r > 0 && nl > 0) {
                ctx->r2 = r; ctx->sweeps2 = 1;
                ctx->sigma_host.assign((size_t)ctx->sig_total, 0.);
                memcpy(ctx->sigma_host.data(), ctx->h_sig, sizeof(double) * (size_t)std::min(ctx->sig_count, ctx->sig_total));
                ctx->sigma_pending = false;
                ctx->n_L = std::min(nl, r); *n_L = ctx->n_L;
                done = true;
            }
        }
    }
    if (!done) {     // larger blocks, a fixed n_L, or a solve that did not end on a clean sweep: the general path
        PB_TRY(pbi_eig_psd(ctx, G2, keep, 0., 0., &ctx->At2, &ctx->At2_cap, ctx->sig2, ctx->nrm2, ctx->perm2,
                           &ctx->r2, &ctx->sweeps2, 40));
        PB_TRY(fetch_sigma(ctx, ctx->sig2, keep, ctx->m));
        PB_TRY(finish_rank(ctx, ctx->sig2, keep, ctx->r2, eps, n_L_in, n_L));
    }
    ctx->pass2_done = true;
    ctx->refine_steps = refine_steps_needed(ctx);
    if (ctx->pass2_t_started) {
        cudaEventRecord(ctx->tev[PB_T_PASS2][1], ctx->stream);
        ctx->t_pending[PB_T_PASS2] = true; ctx->pass2_t_started = false;
    }
    return PB_OK;
}

int pb_pod_basis(pb_ctx* ctx, const double* A, int64_t rows, int64_t lda, const double* Y, double* V, int64_t ldv) {
    PB_ENTER(ctx);
    const int nl = ctx->n_L;
    if (nl <= 0) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod_basis: no decomposition available");
    Timer t(ctx, PB_T_BASIS);
    if (ctx->pass2_done) {
        if (!Y) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod_basis: second-pass modes need the rotated block Y");
        const int keep = ctx->n_keep;
        PB_TRY(pb_reserve(ctx, &ctx->Bmat, &ctx->B_cap, (size_t)keep * nl));
        PB_TRY(pbi_factor_to_right(ctx, ctx->At2, keep, nl, ctx->perm2, ctx->nrm2, ctx->sig2, ctx->Bmat, nl));
        PB_TRY(pbi_gemm_nn(ctx, Y, keep, rows, keep, ctx->Bmat, nl, nl, V, ldv, nullptr, 0, nullptr));
    } else {
        if (ctx->mode != PB_POD_GRAM && ctx->mode != PB_POD_TSQR) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod_basis: second pass not finished");
        const int64_t m = ctx->m;
        PB_TRY(pb_reserve(ctx, &ctx->Bmat, &ctx->B_cap, (size_t)m * nl));
        PB_TRY(pbi_factor_to_right(ctx, ctx->At, m, nl, ctx->perm, ctx->nrm, ctx->sig, ctx->Bmat, nl));
        PB_TRY(pbi_gemm_nn(ctx, A, lda, rows, m, ctx->Bmat, nl, nl, V, ldv, nullptr, 0, nullptr));
    }
    t.stop();
    return PB_OK;
}

int pb_pod_sigma(pb_ctx* ctx, double* sigma_host, int64_t capacity, int64_t* count) {
    PB_ENTER(ctx);
    if (ctx->sigma_pending) {
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->sigma_host.assign((size_t)ctx->sig_total, 0.);
        memcpy(ctx->sigma_host.data(), ctx->h_sig, sizeof(double) * (size_t)std::min(ctx->sig_count, ctx->sig_total));
        ctx->sigma_pending = false;
    }
    int64_t n = std::min<int64_t>(capacity, (int64_t)ctx->sigma_host.size());
    if (sigma_host && n > 0) memcpy(sigma_host, ctx->sigma_host.data(), sizeof(double) * (size_t)n);
    if (count) *count = (int64_t)ctx->sigma_host.size();
    return PB_OK;
}

int pb_pod_right(pb_ctx* ctx, double* Z, int64_t ldz) {
    PB_ENTER(ctx);
    const int nl = ctx->n_L; const int64_t m = ctx->m;
    if (nl <= 0) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod_right: no decomposition available");
    if (ctx->pass2_done) {
        const int keep = ctx->n_keep;
        PB_TRY(pb_reserve(ctx, &ctx->Bmat, &ctx->B_cap, (size_t)keep * nl));
        PB_TRY(pbi_factor_to_right(ctx, ctx->At2, keep, nl, ctx->perm2, ctx->nrm2, nullptr, ctx->Bmat, nl));
        PB_TRY(pbi_gemm_nn(ctx, ctx->Wk, keep, m, keep, ctx->Bmat, nl, nl, Z, ldz, nullptr, 0, nullptr));
    } else {
        PB_TRY(pbi_factor_to_right(ctx, ctx->At, m, nl, ctx->perm, ctx->nrm, nullptr, Z, ldz));
    }
    return PB_OK;
}

// ---------------------------------------------------------------- TSQR mode
int pb_tsqr_r(pb_ctx* ctx, const double* A, int64_t rows, int64_t cols, int64_t lda, double* R) {
    PB_ENTER(ctx); if (!A || !R) return PB_ERR_ARG;
    if (rows <= 0 || cols <= 0 || lda < cols) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_tsqr_r: bad shape (%lld x %lld, ld %lld)",
                                                       (long long)rows, (long long)cols, (long long)lda);
    Timer t(ctx, PB_T_QR);
    ctx->qr_panels = 0; ctx->qr_flops = 0.;
    PB_TRY(pbi_qr_r(ctx, A, rows, cols, lda, R, cols, cols));
    ctx->qr_f_rows = cols;
    t.stop();
    return PB_OK;
}

int pb_tsqr_factor(pb_ctx* ctx, const double* A, int64_t rows, int64_t cols, int64_t lda, double* F, int64_t* f_rows) {
    PB_ENTER(ctx); if (!A || !F || !f_rows) return PB_ERR_ARG;
    if (rows <= 0 || cols <= 0 || lda < cols) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_tsqr_factor: bad shape (%lld x %lld, ld %lld)",
                                                       (long long)rows, (long long)cols, (long long)lda);
    Timer t(ctx, PB_T_QR);
    ctx->qr_panels = 0; ctx->qr_flops = 0.;
    PB_TRY(pbi_qr_factor(ctx, A, rows, cols, lda, F, cols, cols, true, f_rows));
    ctx->qr_f_rows = *f_rows;
    t.stop();
    return PB_OK;
}

int pb_tsqr_stats(pb_ctx* ctx, int64_t* panels, int64_t* f_rows, double* flops) {
    PB_ENTER(ctx);
    if (panels) *panels = ctx->qr_panels;
    if (f_rows) *f_rows = ctx->qr_f_rows;
    if (flops) *flops = ctx->qr_flops;
    return PB_OK;
}

// F (f_rows x m, ld ldf): upper != 0 -> an upper-triangular R (only the triangle is read)
static int pod_from_factor(pb_ctx* ctx, const double* F, int64_t f_rows, int64_t m, int64_t ldf, int upper, double eps,
                           int n_L_in, int* n_L) {
    if (m <= 0 || m > 8192) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_pod_from_r: order %lld outside (0, 8192]", (long long)m);
    if (f_rows <= 0 || f_rows > m || ldf < m) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_pod_from_factor: %lld factor rows for %lld columns (ld %lld)",
                                                       (long long)f_rows, (long long)m, (long long)ldf);
    if (!(eps >= 0.) || eps >= 1.) PB_FAIL(ctx, PB_ERR_ARG, "pb_pod: eps must be in [0, 1)");
    Timer t(ctx, PB_T_EIG);
    ctx->m = m; ctx->mode = PB_POD_TSQR; ctx->pass2_done = false; ctx->n_keep = 0;
    PB_TRY(pb_reserve(ctx, &ctx->sig, &ctx->sig_cap, (size_t)m + 32));
    PB_TRY(pb_reserve(ctx, &ctx->nrm, &ctx->nrm_cap, (size_t)m + 32));
    PB_TRY(pb_reserve_int(ctx, &ctx->perm, &ctx->perm_cap, (size_t)m + 64));
    const int64_t ldm = (m + 7) / 8 * 8, cap_rows = (m + 31) / 32 * 32;
    PB_TRY(pb_reserve(ctx, &ctx->At, &ctx->At_cap, (size_t)cap_rows * ldm));
    if (upper) PB_TRY(pbi_qr_load_factor(ctx, F, ldf, m, ctx->At, ldm, cap_rows));
    else PB_TRY(pbi_qr_load_dense(ctx, F, ldf, f_rows, m, ctx->At, ldm, cap_rows));
    // eps == 0 asks for every mode (pod.py default): no truncation.  Otherwise stop the pivoted QR at the numerical
    // rank: what is left is below m eps |R| -- the backward error of the QR itself
    const bool full = (eps == 0. && n_L_in == 0);
    PB_TRY(pbi_qrcp_factor(ctx, ctx->At, m, f_rows, ldm, cap_rows, full ? 0. : (double)m * DBL_EPSILON, &ctx->r));
    if (ctx->r <= 0) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_pod_from_r: zero matrix");
    PB_TRY(pbi_jacobi_factor(ctx, ctx->At, m, ctx->r, nullptr, ctx->sig, ctx->nrm, ctx->perm, &ctx->sweeps1, 60));
    PB_TRY(fetch_sigma(ctx, ctx->sig, m, m));
    PB_TRY(finish_rank(ctx, ctx->sig, m, ctx->r, eps, n_L_in, n_L));
    t.stop();
    return PB_OK;
}

int pb_pod_from_r(pb_ctx* ctx, const double* R, int64_t m, double eps, int n_L_in, int* n_L) {
    PB_ENTER(ctx); if (!R || !n_L) return PB_ERR_ARG;
    return pod_from_factor(ctx, R, m, m, m, 1, eps, n_L_in, n_L);
}

int pb_pod_from_factor(pb_ctx* ctx, const double* F, int64_t f_rows, int64_t m, int64_t ldf, double eps, int n_L_in, int* n_L) {
    PB_ENTER(ctx); if (!F || !n_L) return PB_ERR_ARG;
    return pod_from_factor(ctx, F, f_rows, m, ldf, 0, eps, n_L_in, n_L);
}

// eigensolve + (optional) second pass on a tall operand whose Gram is already in ctx->Gbuf
static int pod_after_gram(pb_ctx* ctx, const double* A, int64_t rows, int64_t m, int64_t lda, double eps, int n_L_in,
                          int mode, int* n_L) {
    int keep = 0;
    PB_TRY(pb_pod_from_gram(ctx, ctx->Gbuf, m, eps, n_L_in, mode, n_L, &keep));
    ctx->ms[PB_T_PASS2] = 0.; ctx->t_pending[PB_T_PASS2] = false;
    if (keep > 0) {      // (PB_T_PASS2 is armed by the two calls themselves: the sharded drivers issue them separately)
        PB_TRY(pb_reserve(ctx, &ctx->Ybuf, &ctx->Y_cap, (size_t)rows * keep));
        PB_TRY(pb_reserve(ctx, &ctx->G2buf, &ctx->G2_cap, (size_t)keep * keep));
        PB_TRY(pb_pod_pass2_gram(ctx, A, rows, lda, ctx->Ybuf, ctx->G2buf));
        PB_TRY(pb_pod_pass2_finish(ctx, ctx->G2buf, eps, n_L_in, n_L));
    }
    return PB_OK;
}

int pb_pod(pb_ctx* ctx, const double* U, int64_t n_h, int64_t n_st, int64_t ldu, double eps, int n_L_in, int mode, int* n_L) {
    PB_ENTER(ctx); if (!n_L) return PB_ERR_ARG;
    if (n_h <= 0 || n_st <= 0 || ldu < n_st) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_pod: bad shape (%lld x %lld, ld %lld)",
                                                       (long long)n_h, (long long)n_st, (long long)ldu);
    Timer t_total(ctx, PB_T_POD_TOTAL);
    const bool wide = n_h < n_st;
    const double* A = U; int64_t rows = n_h, m = n_st, lda = ldu;
    if (wide) {   // POD of U^T: its right singular vectors are the modes themselves
        PB_TRY(pb_reserve(ctx, &ctx->ATbuf, &ctx->AT_cap, (size_t)n_st * n_h));
        PB_TRY(pbi_transpose(ctx, U, n_h, n_st, ldu, ctx->ATbuf, n_h));
        A = ctx->ATbuf; rows = n_st; m = n_h; lda = n_h;
    }
    if (m > 8192) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_pod: min(n_h, n_st) = %lld exceeds 8192", (long long)m);
    PB_TRY(pb_reserve(ctx, &ctx->Gbuf, &ctx->G_cap, (size_t)m * m));
    ctx->wide = wide;
    if (mode == PB_POD_TSQR || mode == PB_POD_TSQR_FULL) {
        ctx->ms[PB_T_GRAM] = 0.; ctx->ms[PB_T_PASS2] = 0.; ctx->t_pending[PB_T_GRAM] = ctx->t_pending[PB_T_PASS2] = false;
        // eps == 0 with no n_L asks for every mode: nothing may be dropped, so the plain QR runs
        if (mode == PB_POD_TSQR_FULL || (eps == 0. && n_L_in == 0)) {
            PB_TRY(pb_tsqr_r(ctx, A, rows, m, lda, ctx->Gbuf));
            PB_TRY(pb_pod_from_r(ctx, ctx->Gbuf, m, eps, n_L_in, n_L));
        } else {
            int64_t f_rows = 0;
            PB_TRY(pb_tsqr_factor(ctx, A, rows, m, lda, ctx->Gbuf, &f_rows));
            if (f_rows <= 0) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_pod: zero matrix");
            PB_TRY(pb_pod_from_factor(ctx, ctx->Gbuf, f_rows, m, m, eps, n_L_in, n_L));
        }
    } else {
        PB_TRY(pb_gram(ctx, A, rows, m, lda, ctx->Gbuf));
        PB_TRY(pod_after_gram(ctx, A, rows, m, lda, eps, n_L_in, mode, n_L));
    }
    t_total.stop();
    return PB_OK;
}

namespace {
__global__ void scale_kernel(const double* __restrict__ in, int64_t n, double f, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f * in[i];
}
__global__ void sum_partials_kernel(const double* __restrict__ part, int count, int64_t n, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.;
    for (int c = 0; c < count; ++c) s += part[(int64_t)c * n + i];     // fixed order: deterministic
    out[i] = s;
}
}  // namespace

// ---- pageable host buffers: pinned staging ring filled by host threads -------------------------------------------
// cudaMemcpyAsync from pageable memory goes through the driver's own single-threaded bounce buffer (a fraction of the
// PCIe rate).  An ordinary numpy array (what pod.perform_pod is handed) is therefore copied by a few host threads into
// one of PB_STAGE_SLOTS pinned chunk buffers, and the H2D copy of chunk c runs while the threads fill chunk c + 1.
namespace {
bool host_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
int host_copy_threads() {
    static const int n = []{
        const char* e = getenv("PB_HOST_COPY_THREADS");
        int v = e ? atoi(e) : 0;
        if (v <= 0) { v = (int)std::thread::hardware_concurrency() / 2; if (v > 16) v = 16; }
        return v < 1 ? 1 : v;
    }();
    return n;
}
bool host_copy_streaming() {
    static const bool nt = []{ const char* e = getenv("PB_HOST_COPY_NT"); return e ? atoi(e) != 0 : true; }();
    return nt;
}
// one thread's share.  Non-temporal 16-byte stores (PB_HOST_COPY_NT=0 turns them off): no read-for-ownership of the
// destination lines and the staging buffer stays out of the caches.  Alone the plain memcpy is faster on the B200 box's
// host (8 threads: 76 vs 66 GB/s), but with the copy engine reading the ring at the same time the streaming stores win:
// 1.28 GB pageable upload 24.4 ms (52.5 GB/s, the PCIe rate) vs 27.8 ms (profiles/r02_hostcopy.txt)
void copy_span(char* dst, const char* src, size_t bytes, bool nt) {
#if defined(__SSE2__)
    if (nt && bytes >= 4096 && ((uintptr_t)dst & 15) == 0) {
        const size_t n16 = bytes / 64 * 4;
        const __m128i* s = reinterpret_cast<const __m128i*>(src);
        __m128i* d = reinterpret_cast<__m128i*>(dst);
        for (size_t i = 0; i < n16; i += 4) {
            __m128i a = _mm_loadu_si128(s + i), b = _mm_loadu_si128(s + i + 1), c = _mm_loadu_si128(s + i + 2), e = _mm_loadu_si128(s + i + 3);
            _mm_stream_si128(d + i, a); _mm_stream_si128(d + i + 1, b); _mm_stream_si128(d + i + 2, c); _mm_stream_si128(d + i + 3, e);
        }
        _mm_sfence();
        if (bytes > n16 * 16) memcpy(dst + n16 * 16, src + n16 * 16, bytes - n16 * 16);
        return;
    }
#endif
    memcpy(dst, src, bytes);
}
// persistent host threads (started on first use, shared by every context of the process): a job is a row range of a
// pitched copy; the caller takes share 0 itself and waits for the others
class HostCopyPool {
public:
    static HostCopyPool& get() { static HostCopyPool* p = new HostCopyPool(); return *p; }   // never destroyed: no join at exit
    void copy(char* dst, const char* src, size_t rows, size_t row_bytes, size_t src_pitch, int T, bool nt) {
        std::unique_lock<std::mutex> call(call_mu_);            // one copy at a time
        T = std::max(1, std::min(T, kMax));
        ensure(T - 1);
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = {dst, src, rows, row_bytes, src_pitch, T, nt}; pending_ = T - 1; ++gen_;
        }
        if (T > 1) cv_.notify_all();
        run(job_, 0);
        if (T > 1) { std::unique_lock<std::mutex> g(mu_); done_.wait(g, [&]{ return pending_ == 0; }); }
    }
private:
    static constexpr int kMax = 64;
    struct Job { char* dst; const char* src; size_t rows, row_bytes, pitch; int T; bool nt; };
    static void run(const Job& j, int t) {
        if (j.pitch == j.row_bytes) {            // contiguous: 64-byte aligned byte spans
            const size_t total = j.rows * j.row_bytes, lines = total / 64;
            const size_t b0 = lines * t / j.T * 64, b1 = (t + 1 == j.T) ? total : lines * (t + 1) / j.T * 64;
            if (b1 > b0) copy_span(j.dst + b0, j.src + b0, b1 - b0, j.nt);
            return;
        }
        const size_t r0 = j.rows * t / j.T, r1 = j.rows * (t + 1) / j.T;
        for (size_t r = r0; r < r1; ++r) copy_span(j.dst + r * j.row_bytes, j.src + r * j.pitch, j.row_bytes, j.nt);
    }
    void ensure(int n) {
        while ((int)workers_.size() < n) {
            const int id = (int)workers_.size() + 1;
            uint64_t seen; { std::lock_guard<std::mutex> g(mu_); seen = gen_; }
            workers_.emplace_back([this, id, seen]() mutable {
                for (;;) {
                    Job j;
                    { std::unique_lock<std::mutex> g(mu_); cv_.wait(g, [&]{ return gen_ != seen; }); seen = gen_; j = job_; }
                    if (id < j.T) {
                        run(j, id);
                        std::lock_guard<std::mutex> g(mu_);
                        if (--pending_ == 0) done_.notify_one();
                    }
                }
            });
            workers_.back().detach();
        }
    }
    std::mutex call_mu_, mu_; std::condition_variable cv_, done_;
    std::vector<std::thread> workers_; Job job_{}; int pending_ = 0; uint64_t gen_ = 0;
};
// dst (contiguous, row length `row_bytes`) <- src rows with pitch `src_pitch`, split over the host threads
void parallel_host_copy(char* dst, const char* src, size_t rows, size_t row_bytes, size_t src_pitch) {
    const int T = (int)std::min<size_t>((size_t)host_copy_threads(), std::max<size_t>(1, rows * row_bytes / (1u << 20)));
    HostCopyPool::get().copy(dst, src, rows, row_bytes, src_pitch, T, host_copy_streaming());
}
int stage_reserve(pb_ctx* ctx, size_t bytes) {
    if (ctx->stage_cap >= bytes && ctx->stage[0]) return PB_OK;
    for (int s = 0; s < PB_STAGE_SLOTS; ++s) {
        if (ctx->stage[s]) { cudaFreeHost(ctx->stage[s]); ctx->stage[s] = nullptr; }
        cudaError_t e = cudaMallocHost(&ctx->stage[s], bytes);
        if (e != cudaSuccess) { cudaGetLastError(); ctx->stage_cap = 0; PB_FAIL(ctx, PB_ERR_NOMEM, "pinned staging buffer (%zu bytes): %s", bytes, cudaGetErrorString(e)); }
        if (!ctx->stage_ev[s]) PB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[s], cudaEventDisableTiming));
    }
    ctx->stage_cap = bytes;
    return PB_OK;
}
// rows [0, n_rows) of a host matrix (pitch ld_host doubles, n_cols used) -> dev (pitch n_cols) on `stream`.
// pinned source: one async copy; pageable source: through the staging ring (the host thread returns when the last
// chunk has been handed to the copy engine).  `slot` carries the ring position across calls.
int upload_rows(pb_ctx* ctx, double* dev, const double* host, int64_t n_rows, int64_t n_cols, int64_t ld_host, bool pinned,
                cudaStream_t stream, int* slot) {
    const size_t row_bytes = (size_t)n_cols * sizeof(double);
    if (pinned) {
        if (ld_host == n_cols)      // contiguous rows: one linear copy (the pitched path is slower for 8 KB rows)
            PB_CUDA(ctx, cudaMemcpyAsync(dev, host, (size_t)n_rows * row_bytes, cudaMemcpyHostToDevice, stream));
        else
            PB_CUDA(ctx, cudaMemcpy2DAsync(dev, row_bytes, host, ld_host * sizeof(double), row_bytes, n_rows, cudaMemcpyHostToDevice, stream));
        return PB_OK;
    }
    const size_t piece_rows = std::max<size_t>(1, ctx->stage_cap / row_bytes);
    for (int64_t r0 = 0; r0 < n_rows; r0 += (int64_t)piece_rows) {
        const size_t nr = (size_t)std::min<int64_t>((int64_t)piece_rows, n_rows - r0);
        const int s = (*slot)++ % PB_STAGE_SLOTS;
        PB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[s]));          // the copy that last read this slot is done
        parallel_host_copy(reinterpret_cast<char*>(ctx->stage[s]), reinterpret_cast<const char*>(host + r0 * ld_host), nr, row_bytes,
                           (size_t)ld_host * sizeof(double));
        PB_CUDA(ctx, cudaMemcpyAsync(dev + r0 * n_cols, ctx->stage[s], nr * row_bytes, cudaMemcpyHostToDevice, stream));
        PB_CUDA(ctx, cudaEventRecord(ctx->stage_ev[s], stream));
    }
    return PB_OK;
}
int download_staged(pb_ctx* ctx, char* host, const char* dev, size_t bytes) {
    PB_TRY(stage_reserve(ctx, (size_t)64 << 20));
    const size_t piece = ctx->stage_cap;
    const size_t n = (bytes + piece - 1) / piece;
    for (int s = 0; s < PB_STAGE_SLOTS; ++s) PB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[s]));   // ring idle (uploads done)
    auto issue = [&](size_t i) -> cudaError_t {
        const size_t off = i * piece, nb = std::min(piece, bytes - off);
        const int s = (int)(i % PB_STAGE_SLOTS);
        cudaError_t e = cudaMemcpyAsync(ctx->stage[s], dev + off, nb, cudaMemcpyDeviceToHost, ctx->stream);
        return e != cudaSuccess ? e : cudaEventRecord(ctx->stage_ev[s], ctx->stream);
    };
    for (size_t i = 0; i < std::min<size_t>(n, PB_STAGE_SLOTS - 1); ++i) PB_CUDA(ctx, issue(i));
    for (size_t i = 0; i < n; ++i) {
        if (i + PB_STAGE_SLOTS - 1 < n) PB_CUDA(ctx, issue(i + PB_STAGE_SLOTS - 1));   // its slot was drained at step i - 1
        const size_t off = i * piece, nb = std::min(piece, bytes - off);
        const int s = (int)(i % PB_STAGE_SLOTS);
        PB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[s]));
        parallel_host_copy(host + off, static_cast<const char*>(ctx->stage[s]), 1, nb, nb);
    }
    return PB_OK;
}
}  // namespace

// Gram of a HOST matrix: the upload is cut into row chunks on a copy stream and the Gram of chunk c
// runs on the compute stream as soon as chunk c has landed, so the 5 ms Gram hides under the PCIe
// transfer; the chunk Grams are summed in chunk order.  U_dev (n_h x n_st, ld = n_st) receives U.
// Returns 1 when the matrix is too small / wide to chunk (nothing done).
namespace {
int gram_host_chunked(pb_ctx* ctx, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host, double* U_dev,
                      double* G_out) {
    const int64_t chunk_rows_min = 8192;
    int nchunks = (int)std::min<int64_t>(16, n_h / chunk_rows_min);
    if (n_h < n_st || nchunks < 2 || n_st > 8192) return 1;
    const int64_t m = n_st;
    if (!ctx->copy_stream) {
        PB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int c = 0; c < 16; ++c) PB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->chunk_ev[c], cudaEventDisableTiming));
    }
    PB_TRY(pb_reserve(ctx, &ctx->Gpart, &ctx->Gpart_cap, (size_t)m * m * nchunks));
    // the Gram of the LAST chunk cannot overlap anything (nothing is left to upload): keep that chunk small, a quarter of
    // the others (its Gram then takes ~0.1 instead of 0.34 ms at C2)
    int64_t rows_c = (4 * n_h + 4 * nchunks - 4) / (4 * nchunks - 3); rows_c = (rows_c + 15) / 16 * 16;
    // the copy stream must not start overwriting U_dev before earlier work on the compute stream is done
    PB_CUDA(ctx, cudaEventRecord(ctx->chunk_ev[0], ctx->stream));
    PB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[0], 0));
    int used = 0;
    ctx->gram_rows = n_h; ctx->refine_rows_fixed = 0;
    Timer tg(ctx, PB_T_GRAM);
    const bool pinned = host_is_pinned(U_host);
    if (!pinned) PB_TRY(stage_reserve(ctx, (size_t)64 << 20));
    int slot = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int64_t r0 = (int64_t)c * rows_c, nr = std::min(rows_c, n_h - r0);
        if (nr <= 0) break;
        PB_TRY(upload_rows(ctx, U_dev + r0 * n_st, U_host + r0 * ldu_host, nr, n_st, ldu_host, pinned, ctx->copy_stream, &slot));
        PB_CUDA(ctx, cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream));
        PB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[c], 0));
        PB_TRY(pbi_gemm_tn(ctx, U_dev + r0 * n_st, n_st, m, U_dev + r0 * n_st, n_st, m, nr, ctx->Gpart + (size_t)c * m * m, m, 1));
        ++used;
    }
    sum_partials_kernel<<<(unsigned)((m * m + 255) / 256), 256, 0, ctx->stream>>>(ctx->Gpart, used, m * m, G_out);
    PB_LAUNCH_CHECK(ctx);
    tg.stop();
    return PB_OK;
}
}  // namespace

// Rank-revealing TSQR factor of a HOST matrix: the upload is cut into row chunks (an uploader thread fills the copy
// stream: pinned sources directly, pageable ones through the staging ring) and the leaf QR of chunk c runs on the
// compute stream as soon as chunk c has landed -- TSQR leaves in time instead of in space.  The leaf factors
// (f_c x n_st rows each, f_c = a multiple of 64 >= the leaf's numerical rank) are stacked and factored once more, exactly
// as the sharded path does with the factors of the GPUs (test_tsqr_logical_shards_equal_unsharded).  The leaf QR
// synchronises with the host once per panel (pivot decisions), which is why the uploads need their own thread.
// F_out (n_st x n_st) receives the factor, *f_rows its row count.  Returns 1 when the shape does not chunk (nothing done).
namespace {
int tsqr_host_chunked(pb_ctx* ctx, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host, double* U_dev,
                      double* F_out, int64_t* f_rows) {
    static const bool enabled = []{ const char* e = getenv("PB_TSQR_HOST_OVERLAP"); return !e || atoi(e) != 0; }();
    const int64_t m = n_st;
    const int nchunks = (int)std::min<int64_t>(8, n_h / std::max<int64_t>(8192, 8 * m));
    if (!enabled || n_h < n_st || nchunks < 2 || n_st > 8192) return 1;
    if (!ctx->copy_stream) {
        PB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int c = 0; c < 16; ++c) PB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->chunk_ev[c], cudaEventDisableTiming));
    }
    // every leaf zero-fills m rows from its offset (pbi_qr_factor), so the stack keeps one factor of slack
    PB_TRY(pb_reserve(ctx, &ctx->Gpart, &ctx->Gpart_cap, (size_t)(nchunks + 1) * m * m));
    const bool pinned = host_is_pinned(U_host);
    if (!pinned) PB_TRY(stage_reserve(ctx, (size_t)64 << 20));
    int64_t rows_c = (n_h + nchunks - 1) / nchunks; rows_c = (rows_c + 15) / 16 * 16;
    PB_CUDA(ctx, cudaEventRecord(ctx->chunk_ev[15], ctx->stream));      // U_dev may still be read by earlier work
    PB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[15], 0));
    Timer tq(ctx, PB_T_QR);
    ctx->qr_panels = 0; ctx->qr_flops = 0.;
    std::atomic<int> issued{0};
    std::atomic<int> up_rc{PB_OK};
    std::thread uploader([&] {
        cudaSetDevice(ctx->device);
        int slot = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int64_t r0 = (int64_t)c * rows_c, nr = std::min(rows_c, n_h - r0);
            if (nr > 0) {
                int rc = upload_rows(ctx, U_dev + r0 * n_st, U_host + r0 * ldu_host, nr, n_st, ldu_host, pinned, ctx->copy_stream, &slot);
                if (rc == PB_OK && cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream) != cudaSuccess) rc = PB_ERR_CUDA;
                if (rc != PB_OK) { up_rc.store(rc); issued.store(nchunks, std::memory_order_release); return; }
            }
            issued.store(c + 1, std::memory_order_release);
        }
    });
    int rc = PB_OK;
    int64_t off = 0;
    for (int c = 0; c < nchunks && rc == PB_OK; ++c) {
        const int64_t r0 = (int64_t)c * rows_c, nr = std::min(rows_c, n_h - r0);
        if (nr <= 0) break;
        while (issued.load(std::memory_order_acquire) <= c) std::this_thread::yield();
        if (up_rc.load() != PB_OK) { rc = up_rc.load(); break; }
        if (cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[c], 0) != cudaSuccess) { rc = PB_ERR_CUDA; break; }
        int64_t fr = 0;
        rc = pbi_qr_factor(ctx, U_dev + r0 * n_st, nr, m, n_st, ctx->Gpart + (size_t)off * m, m, m, true, &fr);
        off += std::min(fr, m);
    }
    uploader.join();
    if (rc == PB_OK && up_rc.load() != PB_OK) rc = up_rc.load();
    if (rc != PB_OK) return rc < 0 ? rc : PB_ERR_CUDA;
    int64_t fr = 0;
    if (off > 0) PB_TRY(pbi_qr_factor(ctx, ctx->Gpart, off, m, m, F_out, m, m, true, &fr));
    *f_rows = fr; ctx->qr_f_rows = fr;
    tq.stop();
    return PB_OK;
}
}  // namespace

namespace {
// whole-matrix upload on the compute stream (small / wide inputs and the TSQR mode)
int upload_whole(pb_ctx* ctx, double* U_dev, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host) {
    const bool pinned = host_is_pinned(U_host);
    if (!pinned && (size_t)n_h * n_st * sizeof(double) >= ((size_t)8 << 20)) {
        PB_TRY(stage_reserve(ctx, (size_t)64 << 20));
        int slot = 0;
        return upload_rows(ctx, U_dev, U_host, n_h, n_st, ldu_host, false, ctx->stream, &slot);
    }
    PB_CUDA(ctx, cudaMemcpy2DAsync(U_dev, n_st * sizeof(double), U_host, ldu_host * sizeof(double),
                                   n_st * sizeof(double), n_h, cudaMemcpyHostToDevice, ctx->stream));
    return PB_OK;
}
}  // namespace

// Host matrix (n_h x n_st, pitch ldu_host) -> U_dev (pitch n_st) on the context stream, asynchronous with respect to
// the GPU work already queued; pageable sources go through the staging ring (the call returns when the last piece has
// been handed to the copy engine, so the host array may be reused after pb_sync)
int pb_upload_matrix(pb_ctx* ctx, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host, double* U_dev) {
    PB_ENTER(ctx); if (!U_host || !U_dev) return PB_ERR_ARG;
    if (n_h <= 0 || n_st <= 0 || ldu_host < n_st) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_upload_matrix: bad shape");
    return upload_whole(ctx, U_dev, U_host, n_h, n_st, ldu_host);
}

// Host-side ceiling of the staged upload: GB/s of the pageable -> pinned copy alone (threads <= 0: the library's own
// setting; streaming < 0: the library's own).  No GPU work.
int pb_probe_host_copy(size_t bytes, int threads, int streaming, double* gbs) {
    if (!gbs || bytes < (1u << 20)) return PB_ERR_ARG;
    char* src = static_cast<char*>(malloc(bytes));
    void* dst = nullptr;
    if (!src) return PB_ERR_NOMEM;
    const bool pinned = cudaMallocHost(&dst, bytes) == cudaSuccess;
    if (!pinned) { cudaGetLastError(); dst = malloc(bytes); if (!dst) { free(src); return PB_ERR_NOMEM; } memset(dst, 0, bytes); }
    memset(src, 1, bytes);
    const int T = threads > 0 ? threads : host_copy_threads();
    const bool nt = streaming < 0 ? host_copy_streaming() : streaming != 0;
    double best = 0.;
    for (int rep = 0; rep < 4; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        HostCopyPool::get().copy(static_cast<char*>(dst), src, bytes / 4096, 4096, 4096, T, nt);
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rep > 0) best = std::max(best, (double)(bytes / 4096 * 4096) / sec / 1e9);
    }
    if (pinned) cudaFreeHost(dst); else free(dst);
    free(src);
    *gbs = best;
    return PB_OK;
}

int pb_gram_host(pb_ctx* ctx, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host, double* U_dev,
                 double* G_dev) {
    PB_ENTER(ctx); if (!U_host || !U_dev || !G_dev) return PB_ERR_ARG;
    if (n_h <= 0 || n_st <= 0 || ldu_host < n_st) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_gram_host: bad shape");
    int rc = gram_host_chunked(ctx, U_host, n_h, n_st, ldu_host, U_dev, G_dev);
    if (rc != 1) return rc;
    PB_TRY(upload_whole(ctx, U_dev, U_host, n_h, n_st, ldu_host));
    return pb_gram(ctx, U_dev, n_h, n_st, n_st, G_dev);
}

// POD of a HOST matrix (see gram_host_chunked)
int pb_pod_host(pb_ctx* ctx, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host, double* U_dev,
                double eps, int n_L_in, int mode, int* n_L) {
    PB_ENTER(ctx); if (!n_L || !U_host || !U_dev) return PB_ERR_ARG;
    if (n_h <= 0 || n_st <= 0 || ldu_host < n_st) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_pod_host: bad shape");
    Timer t_total(ctx, PB_T_POD_TOTAL);
    const int64_t m = n_st;
    if (n_h >= n_st) PB_TRY(pb_reserve(ctx, &ctx->Gbuf, &ctx->G_cap, (size_t)m * m));
    if (mode == PB_POD_TSQR && n_h >= n_st && !(eps == 0. && n_L_in == 0)) {
        // rank-revealing TSQR: the leaf QR of a row chunk runs while the next chunks cross PCIe
        int64_t f_rows = 0;
        int rq = tsqr_host_chunked(ctx, U_host, n_h, n_st, ldu_host, U_dev, ctx->Gbuf, &f_rows);
        if (rq < 0) return rq;
        if (rq == PB_OK) {
            ctx->wide = false;
            ctx->ms[PB_T_GRAM] = 0.; ctx->ms[PB_T_PASS2] = 0.; ctx->t_pending[PB_T_GRAM] = ctx->t_pending[PB_T_PASS2] = false;
            if (f_rows <= 0) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_pod_host: zero matrix");
            PB_TRY(pb_pod_from_factor(ctx, ctx->Gbuf, f_rows, m, m, eps, n_L_in, n_L));
            t_total.stop();
            return PB_OK;
        }
    }
    // tsqr_full (and shapes that do not chunk) have nothing to overlap with the upload: plain copy, then the device path
    int rc = (mode == PB_POD_TSQR || mode == PB_POD_TSQR_FULL) ? 1 : gram_host_chunked(ctx, U_host, n_h, n_st, ldu_host, U_dev, ctx->Gbuf);
    if (rc == 1) {        // small or wide, or the TSQR mode: upload (staged when pageable), then the device path
        PB_TRY(upload_whole(ctx, U_dev, U_host, n_h, n_st, ldu_host));
        return pb_pod(ctx, U_dev, n_h, n_st, n_st, eps, n_L_in, mode, n_L);
    }
    if (rc != PB_OK) return rc;
    ctx->wide = false;
    PB_TRY(pod_after_gram(ctx, U_dev, n_h, m, n_st, eps, n_L_in, mode, n_L));
    t_total.stop();
    return PB_OK;
}

int pb_pod_basis_full(pb_ctx* ctx, const double* U, int64_t n_h, int64_t n_st, int64_t ldu, double* V, int64_t ldv) {
    PB_ENTER(ctx);
    const int nl = ctx->n_L;
    if (ldv < nl) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_pod_basis_full: ldv %lld < n_L %d", (long long)ldv, nl);
    const bool refine = ctx->pass2_done && ctx->refine_steps > 0;
    if (!ctx->wide) {
        PB_TRY(pb_pod_basis(ctx, U, n_h, ldu, ctx->pass2_done ? ctx->Ybuf : nullptr, V, ldv));
        if (!refine) return PB_OK;
        Timer t(ctx, PB_T_REFINE);
        PB_TRY(pb_reserve(ctx, &ctx->Pbuf, &ctx->P_cap, (size_t)n_st * nl + (size_t)nl * nl));
        double* B = ctx->Pbuf; double* G3 = B + (size_t)n_st * nl;
        PB_TRY(pb_project(ctx, V, ldv, nl, U, ldu, n_h, n_st, B));          // B = (V^T U)^T = U^T V  (n_st x n_L)
        PB_TRY(pb_pod_refine_apply(ctx, U, n_h, ldu, B, G3));
        PB_TRY(pb_pod_refine_finish(ctx, G3, n_h, V, ldv));
        t.stop();
        return PB_OK;
    }
    Timer t(ctx, PB_T_BASIS);
    PB_TRY(pb_pod_right(ctx, V, ldv));        // wide: the modes are the right factor of U^T
    t.stop();
    if (refine) {
        // wide case: A = U^T (n_st x n_h, in ctx->ATbuf) is the tall operand, the modes are its RIGHT singular vectors:
        // Z~ = A^T (A Z) Sigma^-2 = U (U^T Z) Sigma^-2, then the Cholesky QR of the n_h x n_L block
        Timer t2(ctx, PB_T_REFINE);
        const double* A = ctx->ATbuf; const int64_t rows = n_st, m = n_h;
        PB_TRY(pb_reserve(ctx, &ctx->Pbuf, &ctx->P_cap, (size_t)rows * nl + (size_t)nl * nl));
        double* B = ctx->Pbuf; double* G3 = B + (size_t)rows * nl;
        PB_TRY(pbi_gemm_nn(ctx, A, m, rows, m, V, ldv, nl, B, nl, nullptr, 0, nullptr));                 // B = A Z (n_st x n_L)
        scale_cols_inv_sq_kernel<<<(unsigned)((rows * nl + 255) / 256), 256, 0, ctx->stream>>>(B, rows, nl, ctx->sig2);
        PB_LAUNCH_CHECK(ctx);
        PB_TRY(pb_reserve(ctx, &ctx->Rbuf, &ctx->R_cap, (size_t)m * nl));
        PB_TRY(pbi_gemm_tn(ctx, A, m, m, B, nl, nl, rows, ctx->Rbuf, nl, 0));                              // Z~ = A^T B (n_h x n_L)
        PB_TRY(pbi_gemm_tn(ctx, ctx->Rbuf, nl, nl, ctx->Rbuf, nl, nl, m, G3, nl, 1));
        PB_TRY(pb_pod_refine_finish(ctx, G3, m, V, ldv));
        t2.stop();
    }
    return PB_OK;
}

// ------------------------------------------------- two-step POD (pod.py:51-68)
namespace {
// copy the first r_z columns of T_z (rows x m) into U_f at column offset off_z
__global__ void pack_bases_kernel(const double* __restrict__ T, int64_t rows, int m, const int* __restrict__ ranks,
                                  const int* __restrict__ offs, double* __restrict__ Uf, int64_t ldf) {
    const int z = blockIdx.y;
    const int r = ranks[z]; const int64_t off = offs[z];
    const double* Tz = T + (int64_t)z * rows * m;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < rows * r; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t h = idx / r; const int c = (int)(idx % r);
        Uf[h * ldf + off + c] = Tz[h * m + c];
    }
}

// Stage 1 of the two-step POD as ONE batch: every kernel below is launched once for all samples.
int fast_pod_stage1_batched(pb_ctx* ctx, const double* U, int64_t n_h, int n_t, int64_t n_s, int64_t ldu,
                            double eps_init, int mode, int* ranks, int64_t ldf, int64_t* total,
                            const std::function<int(double*, int64_t)>* reduce = nullptr) {
    const int64_t m = n_t; const int B = (int)n_s;
    const int64_t ldm = (m + 7) / 8 * 8, cap_rows = (m + 31) / 32 * 32;
    enum { G = 0, AT, SIG, NRM, SCR, BM, T1, T2 };
    PB_TRY(pb_reserve(ctx, &ctx->fp_buf[G], &ctx->fp_cap[G], (size_t)B * m * m));
    PB_TRY(pb_reserve(ctx, &ctx->fp_buf[AT], &ctx->fp_cap[AT], (size_t)B * cap_rows * ldm));
    PB_TRY(pb_reserve(ctx, &ctx->fp_buf[SIG], &ctx->fp_cap[SIG], (size_t)B * m));
    PB_TRY(pb_reserve(ctx, &ctx->fp_buf[NRM], &ctx->fp_cap[NRM], (size_t)B * m));
    PB_TRY(pb_reserve(ctx, &ctx->fp_buf[SCR], &ctx->fp_cap[SCR], (size_t)B * (cap_rows + 2) + 16));
    PB_TRY(pb_reserve(ctx, &ctx->fp_buf[BM], &ctx->fp_cap[BM], (size_t)B * m * m));
    PB_TRY(pb_reserve(ctx, &ctx->fp_buf[T1], &ctx->fp_cap[T1], (size_t)B * n_h * m));
    PB_TRY(pb_reserve_int(ctx, &ctx->fp_ibuf, &ctx->fp_icap, (size_t)B * (cap_rows + 4)));
    double *Gall = ctx->fp_buf[G], *At = ctx->fp_buf[AT], *sig = ctx->fp_buf[SIG], *nrm = ctx->fp_buf[NRM],
           *scr = ctx->fp_buf[SCR], *Bm = ctx->fp_buf[BM], *Tm = ctx->fp_buf[T1];
    int *perm = ctx->fp_ibuf, *r_dev = perm + (size_t)B * cap_rows, *nl_dev = r_dev + B, *off_dev = nl_dev + B;
    const bool full = (mode == PB_POD_ACCURATE) || eps_init == 0.;
    PB_TRY(pbi_gram_batched(ctx, U, ldu, m, n_t, n_h, Gall, m * m, B));
    if (reduce) PB_TRY((*reduce)(Gall, (int64_t)B * m * m));          // row-sharded: sum of the per-GPU partial Grams
    PB_TRY(pbi_eig_psd_batched(ctx, Gall, m, B, full ? 0. : 4. * DBL_EPSILON, full ? 2. * (double)m * DBL_EPSILON : 0.,
                               At, sig, nrm, perm, r_dev, scr, nullptr, 40));
    const double* basis_src = Tm;
    if (mode == PB_POD_GRAM) {
        PB_TRY(pbi_rank_batched(ctx, sig, m, B, eps_init, nl_dev));
        PB_TRY(pbi_factor_to_right_batched(ctx, At, m, (int)m, B, perm, nrm, sig, Bm, m * m));
        PB_TRY(pbi_gemm_nn_batched(ctx, U, ldu, n_t, n_h, m, Bm, m, m * m, m, Tm, m, n_h * m, B));
    } else {
        // second pass per sample: Y = U_k W_k, G2 = Y^T Y, Jacobi again, T_k = Y W2 / sigma
        PB_TRY(pb_reserve(ctx, &ctx->fp_buf[T2], &ctx->fp_cap[T2], (size_t)B * n_h * m));
        double* T2m = ctx->fp_buf[T2];
        PB_TRY(pbi_factor_to_right_batched(ctx, At, m, (int)m, B, perm, nrm, nullptr, Bm, m * m));
        PB_TRY(pbi_gemm_nn_batched(ctx, U, ldu, n_t, n_h, m, Bm, m, m * m, m, Tm, m, n_h * m, B));       // Y
        PB_TRY(pbi_gram_batched(ctx, Tm, m, m, n_h * m, n_h, Gall, m * m, B));                            // G2
        if (reduce) PB_TRY((*reduce)(Gall, (int64_t)B * m * m));
        PB_TRY(pbi_eig_psd_batched(ctx, Gall, m, B, 0., 0., At, sig, nrm, perm, r_dev, scr, nullptr, 40));
        PB_TRY(pbi_rank_batched(ctx, sig, m, B, eps_init, nl_dev));
        PB_TRY(pbi_factor_to_right_batched(ctx, At, m, (int)m, B, perm, nrm, sig, Bm, m * m));
        PB_TRY(pbi_gemm_nn_batched(ctx, Tm, m, n_h * m, n_h, m, Bm, m, m * m, m, T2m, m, n_h * m, B));
        basis_src = T2m;
    }
    // ranks to the host (one sync), clamp to the numerical rank, prefix offsets back to the device
    std::vector<int> nl(B), rr(B), off(B);
    PB_CUDA(ctx, cudaMemcpyAsync(nl.data(), nl_dev, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(rr.data(), r_dev, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t acc = 0;
    for (int z = 0; z < B; ++z) {
        if (rr[z] <= 0) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_fast_pod: sample %d has no positive pivot (zero trajectory)", z);
        nl[z] = std::min(nl[z], rr[z]);
        off[z] = (int)acc; acc += nl[z];
        if (ranks) ranks[z] = nl[z];
    }
    if (acc > ldf) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_fast_pod: packed basis overflow");
    PB_CUDA(ctx, cudaMemcpyAsync(nl_dev, nl.data(), sizeof(int) * B, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(off_dev, off.data(), sizeof(int) * B, cudaMemcpyHostToDevice, ctx->stream));
    pack_bases_kernel<<<dim3(64, B), 256, 0, ctx->stream>>>(basis_src, n_h, (int)m, nl_dev, off_dev, ctx->Uf, ldf);
    PB_LAUNCH_CHECK(ctx);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // host vectors go out of scope
    *total = acc;
    return PB_OK;
}
}  // namespace

}  // extern "C"
// Row-sharded stage 1 (multi.cu): the local rows of every sample's basis, packed into ctx->Uf (n_h_loc x ldf, ldf =
// n_t n_s); `reduce` sums the batched Gram partials over the GPUs, the per-sample eigensolves are replicated.
int pbi_fast_pod_stage1(pb_ctx* ctx, const double* U, int64_t n_h_loc, int n_t, int64_t n_s, int64_t ldu, double eps_init,
                        int mode, int* ranks_host, int64_t* total, const std::function<int(double*, int64_t)>& reduce) {
    const int64_t ldf = (int64_t)n_t * n_s;
    PB_TRY(pb_reserve(ctx, &ctx->Uf, &ctx->Uf_cap, (size_t)std::max<int64_t>(n_h_loc, 1) * ldf));
    PB_TRY(fast_pod_stage1_batched(ctx, U, n_h_loc, n_t, n_s, ldu, eps_init, mode == PB_POD_TSQR ? PB_POD_ACCURATE : mode,
                                   ranks_host, ldf, total, &reduce));
    ctx->uf_rows = n_h_loc; ctx->uf_cols = *total; ctx->uf_ld = ldf;
    return PB_OK;
}
extern "C" {

int pb_fast_pod(pb_ctx* ctx, const double* U, int64_t n_h, int n_t, int64_t n_s, int64_t ldu,
                double eps, double eps_init, int mode, int* n_L, int* ranks_host) {
    PB_ENTER(ctx); if (!n_L) return PB_ERR_ARG;
    if (n_t <= 0 || n_s <= 0 || ldu < n_s * n_t) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_fast_pod: bad shape");
    if (mode < PB_POD_GRAM || mode > PB_POD_TSQR) PB_FAIL(ctx, PB_ERR_ARG, "pb_fast_pod: unknown mode %d", mode);
    if (!(eps >= 0.) || eps >= 1. || !(eps_init >= 0.) || eps_init >= 1.) PB_FAIL(ctx, PB_ERR_ARG, "pb_fast_pod: eps in [0, 1)");
    cudaEvent_t e0 = ctx->ev2, e1 = ctx->ev3;          // context-owned (the inner pb_pod calls use their own pairs)
    cudaEventRecord(e0, ctx->stream);
    // stage 1: per-sample bases T_k, packed side by side into U_f (n_h x sum r_k); r_k <= min(n_h, n_t)
    const int64_t cap_per = std::min<int64_t>(n_h, n_t);
    const int64_t ldf = cap_per * n_s;
    int rc = pb_reserve(ctx, &ctx->Uf, &ctx->Uf_cap, (size_t)n_h * ldf);
    int64_t off = 0;
    static const bool batched_ok = []{ const char* e = getenv("PB_FASTPOD_BATCHED"); return !e || atoi(e) != 0; }();
    if (rc == PB_OK && batched_ok && n_h >= n_t && n_t <= 1024 && n_s < 65536) {
        // TSQR mode: the per-sample stage runs the batched accurate path (n_s small PODs of n_h x n_t), the global
        // stage below the TSQR path
        rc = fast_pod_stage1_batched(ctx, U, n_h, n_t, n_s, ldu, eps_init, mode == PB_POD_TSQR ? PB_POD_ACCURATE : mode,
                                     ranks_host, ldf, &off);
    } else {
        for (int64_t k = 0; k < n_s && rc == PB_OK; ++k) {     // general shapes: one POD per sample
            int rk = 0;
            const double* Uk = U + k * n_t;
            rc = pb_pod(ctx, Uk, n_h, n_t, ldu, eps_init, 0, mode, &rk);
            if (rc == PB_OK) rc = pb_pod_basis_full(ctx, Uk, n_h, n_t, ldu, ctx->Uf + off, ldf);
            if (ranks_host) ranks_host[k] = rk;
            off += rk;
        }
    }
    // stage 2: global POD of the unweighted concatenation (usually wide: n_h < sum r_k)
    if (rc == PB_OK) rc = pb_pod(ctx, ctx->Uf, n_h, off, ldf, eps, 0, mode, n_L);
    ctx->uf_rows = n_h; ctx->uf_cols = off; ctx->uf_ld = ldf;
    cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ctx->ms[PB_T_POD_TOTAL] = ms; ctx->t_pending[PB_T_POD_TOTAL] = false;
    return rc;
}

int pb_fast_pod_basis(pb_ctx* ctx, double* V, int64_t ldv) {
    PB_ENTER(ctx);
    if (!ctx->Uf || ctx->uf_cols <= 0) PB_FAIL(ctx, PB_ERR_ARG, "pb_fast_pod_basis: call pb_fast_pod first");
    return pb_pod_basis_full(ctx, ctx->Uf, ctx->uf_rows, ctx->uf_cols, ctx->uf_ld, V, ldv);
}

namespace {
// (n_h, n_t, n_s) C-order  <->  (n_h, n_s * n_t) with column i * n_t + j
__global__ void struct_permute_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t n_h,
                                      int n_t, int64_t n_s, int64_t ldu, int to_matrix) {
    __shared__ double tile[32][33];
    // per row h: an (n_t x n_s) <-> (n_s x n_t) transpose
    const int64_t h = blockIdx.z;
    const int64_t s0 = (int64_t)blockIdx.x * 32; const int t0 = blockIdx.y * 32;
    if (to_matrix) {
        for (int y = threadIdx.y; y < 32; y += blockDim.y) {
            int t = t0 + y; int64_t s = s0 + threadIdx.x;
            tile[y][threadIdx.x] = (t < n_t && s < n_s) ? in[(h * n_t + t) * n_s + s] : 0.;
        }
        __syncthreads();
        for (int y = threadIdx.y; y < 32; y += blockDim.y) {
            int64_t s = s0 + y; int t = t0 + threadIdx.x;
            if (s < n_s && t < n_t) out[h * ldu + s * n_t + t] = tile[threadIdx.x][y];
        }
    } else {
        for (int y = threadIdx.y; y < 32; y += blockDim.y) {
            int64_t s = s0 + y; int t = t0 + threadIdx.x;
            tile[y][threadIdx.x] = (s < n_s && t < n_t) ? in[h * ldu + s * n_t + t] : 0.;
        }
        __syncthreads();
        for (int y = threadIdx.y; y < 32; y += blockDim.y) {
            int t = t0 + y; int64_t s = s0 + threadIdx.x;
            if (t < n_t && s < n_s) out[(h * n_t + t) * n_s + s] = tile[threadIdx.x][y];
        }
    }
}
int struct_permute(pb_ctx* ctx, const double* in, double* out, int64_t n_h, int n_t, int64_t n_s, int64_t ldu, int to_matrix) {
    if (n_h <= 0 || n_t <= 0 || n_s <= 0 || ldu < n_s * n_t) PB_FAIL(ctx, PB_ERR_SHAPE, "struct permute: bad shape");
    for (int64_t h0 = 0; h0 < n_h; h0 += 65535) {       // grid.z limit
        int64_t nh = std::min<int64_t>(65535, n_h - h0);
        dim3 grid((unsigned)((n_s + 31) / 32), (unsigned)((n_t + 31) / 32), (unsigned)nh), block(32, 8);
        const double* i2 = to_matrix ? in + h0 * n_t * n_s : in + h0 * ldu;
        double* o2 = to_matrix ? out + h0 * ldu : out + h0 * n_t * n_s;
        struct_permute_kernel<<<grid, block, 0, ctx->stream>>>(i2, o2, nh, n_t, n_s, ldu, to_matrix);
        PB_LAUNCH_CHECK(ctx);
    }
    return PB_OK;
}
}  // namespace

int pb_struct_to_matrix(pb_ctx* ctx, const double* Us, int64_t n_h, int n_t, int64_t n_s, double* U, int64_t ldu) {
    PB_ENTER(ctx);
    return struct_permute(ctx, Us, U, n_h, n_t, n_s, ldu, 1);
}
int pb_matrix_to_struct(pb_ctx* ctx, const double* U, int64_t ldu, int64_t n_h, int n_t, int64_t n_s, double* Us) {
    PB_ENTER(ctx);
    return struct_permute(ctx, U, Us, n_h, n_t, n_s, ldu, 0);
}

// ------------------------------------------------- projection / reconstruction
int pb_project(pb_ctx* ctx, const double* V, int64_t ldv, int n_L, const double* U, int64_t ldu,
               int64_t rows, int64_t n, double* v) {
    PB_ENTER(ctx);
    if (n_L <= 0 || n <= 0 || ldv < n_L || ldu < n) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_project: bad shape");
    Timer t(ctx, PB_T_PROJECT);
    // C (n_L x n) = V^T U, then v = C^T (the reference returns the transposed view, podnnmodel.py:437)
    PB_TRY(pb_reserve(ctx, &ctx->Bmat, &ctx->B_cap, (size_t)n_L * n));
    PB_TRY(pbi_gemm_tn(ctx, V, ldv, n_L, U, ldu, n, rows, ctx->Bmat, n, 0));
    PB_TRY(pbi_transpose(ctx, ctx->Bmat, n_L, n, n, v, n_L));
    t.stop();
    return PB_OK;
}

int pb_reconstruct(pb_ctx* ctx, const double* V, int64_t ldv, int n_L, const double* v, int64_t n, int64_t rows,
                   const double* U, int64_t ldu, double* U_pod, int64_t ldp, double* pod_sig) {
    PB_ENTER(ctx);
    if (n_L <= 0 || n <= 0 || ldv < n_L) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_reconstruct: bad shape");
    if (pod_sig && !U) PB_FAIL(ctx, PB_ERR_ARG, "pb_reconstruct: pod_sig needs U");
    if (!U_pod && !pod_sig) PB_FAIL(ctx, PB_ERR_ARG, "pb_reconstruct: nothing to compute");
    Timer t(ctx, PB_T_RECON);
    // B = v^T (n_L x n), padded to an even leading dimension for 16-byte copies
    const int64_t ldb = (n + 1) / 2 * 2;
    PB_TRY(pb_reserve(ctx, &ctx->Bmat, &ctx->B_cap, (size_t)n_L * ldb));
    PB_TRY(pbi_transpose(ctx, v, n, n_L, n_L, ctx->Bmat, ldb));
    PB_TRY(pbi_gemm_nn(ctx, V, ldv, rows, n_L, ctx->Bmat, ldb, n, U_pod, ldp, pod_sig ? U : nullptr, ldu, pod_sig));
    t.stop();
    return PB_OK;
}

// ---- measurement helper: residual of a computed basis against a reference basis ------------------------------
// R = V - Phi C over the local rows, G_out = R^T R (n_L x n_L).  With C = Phi^T V summed over all rows (pb_gemm_tn +
// all-reduce) and Phi orthonormal, sqrt(lambda_max(sum over ranks of G_out)) = sin of the largest principal angle
// between span(V) and span(Phi) -- computed from the residual itself, so angles of 1e-12 are resolved (1 - cos is not).
int pb_residual_gram(pb_ctx* ctx, const double* V, int64_t ldv, int n_L, const double* Phi, int64_t ldp, int k,
                     int64_t rows, const double* C_dev, double* G_out) {
    PB_ENTER(ctx); if (!V || !Phi || !C_dev || !G_out) return PB_ERR_ARG;
    if (n_L <= 0 || k <= 0 || rows <= 0 || ldv < n_L || ldp < k) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_residual_gram: bad shape");
    // own scratch (a measurement helper must not disturb the buffers of the last decomposition)
    double* R = nullptr; double* Cn = nullptr;
    PB_CUDA(ctx, cudaMalloc(&R, sizeof(double) * (size_t)rows * n_L));
    if (cudaMalloc(&Cn, sizeof(double) * (size_t)k * n_L) != cudaSuccess) { cudaGetLastError(); cudaFree(R); PB_FAIL(ctx, PB_ERR_NOMEM, "pb_residual_gram: out of memory"); }
    int rc = PB_OK;
    if (cudaMemcpy2DAsync(R, n_L * sizeof(double), V, ldv * sizeof(double), n_L * sizeof(double), rows,
                          cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) rc = PB_ERR_CUDA;
    scale_kernel<<<(unsigned)(((int64_t)k * n_L + 255) / 256), 256, 0, ctx->stream>>>(C_dev, (int64_t)k * n_L, -1., Cn);
    ctx->launches++;
    if (rc == PB_OK) rc = pbi_gemm_nn_sub(ctx, Phi, ldp, rows, k, Cn, n_L, n_L, R, n_L);      // R = V + Phi (-C)
    if (rc == PB_OK) rc = pbi_gemm_tn(ctx, R, n_L, n_L, R, n_L, n_L, rows, G_out, n_L, 0);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(R); cudaFree(Cn);
    return rc;
}

}  // extern "C"
