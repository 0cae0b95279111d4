// Single-process multi-GPU POD (pb_multi_*): one context, one stream and one host thread per GPU, the snapshot matrix
// row-sharded by dof, the exchange steps of the path (sum of the n_st x n_st Gram partials, of the second-pass
// Gram, of the projection partials; TSQR mode: gather of the per-GPU factors) done by the library's own kernels over
// NVLink peer memory -- no MPI, no torchrun, no NCCL, no torch.  The reference has no multi-GPU path of its own
// (Horovod carries a rank id only, experiments/*/train.py:19-33); its single-process call pod.perform_pod(U, eps)
// (pod.py:7-47) is what pb_multi_pod_host stands behind when n_gpus > 1.
//
// All-reduce (deterministic, every GPU ends with the same bits): buffer cut into P slices; GPU p sums slice p of all P
// buffers in rank order through peer loads and writes it back into its own buffer (reduce-scatter), then copies the
// other P - 1 reduced slices from their owners (all-gather, copy engine over NVLink).  The three hand-offs are
// stream events exchanged between the per-GPU threads at host barriers; nothing spins on the device.
#include "common.cuh"
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <algorithm>

namespace {

constexpr int MAX_GPUS = 16;

struct PeerPtrs { const double* p[MAX_GPUS]; };

// out[i] = sum_q in.p[q][i] for i in [lo, hi), rank order (the same order on every GPU -> identical bits everywhere)
__global__ void __launch_bounds__(256) peer_reduce_kernel(PeerPtrs in, int P, double* __restrict__ out, int64_t lo, int64_t hi) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool vec = ((lo | hi) & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (int q = 0; q < P; ++q) vec = vec && (reinterpret_cast<uintptr_t>(in.p[q]) & 15) == 0;
    if (vec) {
        for (int64_t j = lo / 2 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < hi / 2; j += stride) {
            double2 s = reinterpret_cast<const double2*>(in.p[0])[j];
            for (int q = 1; q < P; ++q) { const double2 v = reinterpret_cast<const double2*>(in.p[q])[j]; s.x += v.x; s.y += v.y; }
            reinterpret_cast<double2*>(out)[j] = s;
        }
        return;
    }
    for (; i < hi; i += stride) {
        double s = in.p[0][i];
        for (int q = 1; q < P; ++q) s += in.p[q][i];
        out[i] = s;
    }
}

class SpinBarrier {
public:
    explicit SpinBarrier(int n) : n_(n) {}
    void wait() {
        const int gen = gen_.load(std::memory_order_acquire);
        if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            count_.store(0, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
            return;
        }
        int spins = 0;
        while (gen_.load(std::memory_order_acquire) == gen) { if (++spins > 2000) std::this_thread::yield(); }
    }
private:
    const int n_; std::atomic<int> count_{0}; std::atomic<int> gen_{0};
};

}  // namespace

struct pb_multi {
    int P = 0;
    std::vector<int> dev;
    std::vector<pb_ctx*> ctx;
    std::string err;
    // workers
    std::vector<std::thread> th;
    std::mutex mu; std::condition_variable cv, done_cv;
    std::function<int(int)> job; uint64_t gen = 0; int pending = 0; bool quit = false;
    std::vector<int> rc;
    SpinBarrier* bar = nullptr;
    std::vector<int> step_rc;                      // agree(): per-rank status of the step before a collective
    std::vector<cudaEvent_t> ev[3];                // ready / reduced / done events, one per GPU
    std::vector<const double*> xptr;               // exchange: the buffer every rank contributes to the running collective
    std::vector<int64_t> xval;                     // exchange: small host integers (rows of the truncated factors)
    // ---- state of the last decomposition
    std::vector<const double*> U; std::vector<int64_t> rows, ldu;
    int64_t n_h = 0, n_st = 0; int n_L = 0, mode = 0, keep = 0; bool have = false, host_slabs = false;
    std::vector<double*> V; std::vector<size_t> V_cap; std::vector<int64_t> ldv;   // library-owned basis slabs (host-facing calls)
    std::vector<double*> F; std::vector<size_t> F_cap;                              // TSQR: factor / stack buffers
    std::vector<double*> S; std::vector<size_t> S_cap;
    std::vector<double*> vb; std::vector<size_t> vb_cap;                            // projection result, replicated
    std::vector<cudaEvent_t> t0, t1;
};

namespace {

int run(pb_multi* m, std::function<int(int)> f) {
    {
        std::lock_guard<std::mutex> g(m->mu);
        m->job = std::move(f); m->pending = m->P; ++m->gen;
        std::fill(m->rc.begin(), m->rc.end(), PB_OK);
    }
    m->cv.notify_all();
    std::unique_lock<std::mutex> g(m->mu);
    m->done_cv.wait(g, [&]{ return m->pending == 0; });
    for (int p = 0; p < m->P; ++p)
        if (m->rc[p] != PB_OK) { m->err = "GPU " + std::to_string(m->dev[p]) + " (rank " + std::to_string(p) + "): " + m->ctx[p]->err; return m->rc[p]; }
    return PB_OK;
}

void worker(pb_multi* m, int p) {
    cudaSetDevice(m->dev[p]);
    uint64_t seen = 0;
    for (;;) {
        std::function<int(int)> f;
        {
            std::unique_lock<std::mutex> g(m->mu);
            m->cv.wait(g, [&]{ return m->quit || m->gen != seen; });
            if (m->quit) return;
            seen = m->gen; f = m->job;
        }
        const int rc = f(p);
        std::lock_guard<std::mutex> g(m->mu);
        m->rc[p] = rc;
        if (--m->pending == 0) m->done_cv.notify_all();
    }
}

// Every rank reports the status of the step it just finished; all ranks get the first failure (so that nobody enters a
// collective a peer will never join).
int agree(pb_multi* m, int p, int rc) {
    m->step_rc[p] = rc;
    m->bar->wait();
    int out = PB_OK;
    for (int q = 0; q < m->P; ++q) if (m->step_rc[q] != PB_OK) { out = m->step_rc[q]; if (q != p && rc == PB_OK) m->ctx[p]->err = "aborted: " + m->ctx[q]->err; break; }
    m->bar->wait();
    return out;
}
#define MB_STEP(expr) do { int _r = agree(m, p, (expr)); if (_r != PB_OK) return _r; } while (0)

// hand-off: every rank records ev[k][p] on its stream; after the host barrier every rank's stream waits for all of them
int exchange_events(pb_multi* m, int p, int k) {
    pb_ctx* c = m->ctx[p];
    cudaError_t e = cudaEventRecord(m->ev[k][p], c->stream);
    m->step_rc[p] = e == cudaSuccess ? PB_OK : PB_ERR_CUDA;
    m->bar->wait();
    for (int q = 0; q < m->P; ++q)
        if (q != p && e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, m->ev[k][q], 0);
    m->bar->wait();          // nobody re-records ev[k] before every peer has queued its wait
    if (e != cudaSuccess) { c->err = std::string("multi-GPU event exchange: ") + cudaGetErrorString(e); cudaGetLastError(); return PB_ERR_CUDA; }
    return PB_OK;
}

// in-place sum of buf (count doubles) over the ranks; every rank returns with the same bits
int all_reduce(pb_multi* m, int p, double* buf, int64_t count) {
    const int P = m->P;
    if (P == 1) return PB_OK;
    pb_ctx* c = m->ctx[p];
    m->xptr[p] = buf;
    PB_TRY(exchange_events(m, p, 0));                       // all partials are complete (and xptr is published)
    PeerPtrs in;
    for (int q = 0; q < P; ++q) in.p[q] = m->xptr[q];
    const int64_t per = ((count + P - 1) / P + 1) / 2 * 2;  // even slice length: 16-byte accesses
    const int64_t lo = std::min<int64_t>(count, per * p), hi = std::min<int64_t>(count, per * (p + 1));
    if (hi > lo) {
        const int64_t work = (hi - lo + 1) / 2;
        const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)c->sm_count * 8));
        peer_reduce_kernel<<<blocks, 256, 0, c->stream>>>(in, P, buf, lo, hi);
        PB_LAUNCH_CHECK(c);
    }
    PB_TRY(exchange_events(m, p, 1));                       // every slice is reduced at its owner
    for (int q = 0; q < P; ++q) {
        if (q == p) continue;
        const int64_t qlo = std::min<int64_t>(count, per * q), qhi = std::min<int64_t>(count, per * (q + 1));
        if (qhi > qlo)
            PB_CUDA(c, cudaMemcpyPeerAsync(buf + qlo, m->dev[p], m->xptr[q] + qlo, m->dev[q], (size_t)(qhi - qlo) * sizeof(double), c->stream));
    }
    PB_TRY(exchange_events(m, p, 2));                       // peers are done reading this rank's slice
    return PB_OK;
}

// stack (P * blk doubles) <- rank q's block at offset q * blk
int all_gather(pb_multi* m, int p, const double* mine, int64_t blk, double* stack) {
    const int P = m->P;
    pb_ctx* c = m->ctx[p];
    m->xptr[p] = mine;
    PB_TRY(exchange_events(m, p, 0));
    for (int q = 0; q < P; ++q)
        PB_CUDA(c, cudaMemcpyPeerAsync(stack + (int64_t)q * blk, m->dev[p], m->xptr[q], m->dev[q], (size_t)blk * sizeof(double), c->stream));
    PB_TRY(exchange_events(m, p, 2));
    return PB_OK;
}

int reserve(pb_ctx* c, std::vector<double*>& v, std::vector<size_t>& cap, int p, size_t elems) {
    return pb_reserve(c, &v[p], &cap[p], std::max<size_t>(elems, 1));
}

void row_range(int64_t n, int p, int P, int64_t* lo, int64_t* hi) {
    const int64_t base = n / P, rem = n % P;
    *lo = p * base + std::min<int64_t>(p, rem);
    *hi = *lo + base + (p < rem ? 1 : 0);
}

// ---- the sharded decomposition, run by rank p (slab already in HBM unless U_host != nullptr) -------------------------
int pod_rank(pb_multi* m, int p, const double* U_host, int64_t ldu_host, int64_t row0, double eps, int n_L_in, int mode) {
    pb_ctx* c = m->ctx[p];
    const int64_t rows = m->rows[p], n = m->n_st, ld = m->ldu[p];
    double* Udev = const_cast<double*>(m->U[p]);
    int nl = 0;
    if (mode == PB_POD_TSQR || mode == PB_POD_TSQR_FULL) {
        int rc = PB_OK;
        if (U_host && rows > 0) rc = pb_upload_matrix(c, U_host + row0 * ldu_host, rows, n, ldu_host, Udev);
        MB_STEP(rc);
        const bool plain = mode == PB_POD_TSQR_FULL || (eps == 0. && n_L_in == 0);
        MB_STEP(reserve(c, m->F, m->F_cap, p, (size_t)n * n));
        int64_t fr = 0;
        if (rows > 0) rc = plain ? pb_tsqr_r(c, Udev, rows, n, ld, m->F[p]) : pb_tsqr_factor(c, Udev, rows, n, ld, m->F[p], &fr);
        else rc = pb_memset(c, m->F[p], 0, (size_t)n * n * sizeof(double));
        if (plain) fr = rows > 0 ? n : 0;
        m->xval[p] = fr;
        MB_STEP(rc);                                        // (the barrier inside publishes xval)
        int64_t cap = 0;
        for (int q = 0; q < m->P; ++q) cap = std::max(cap, m->xval[q]);
        if (cap <= 0) { c->err = "pb_multi_pod: zero matrix"; MB_STEP(PB_ERR_SHAPE); }
        const double* F = m->F[p]; int64_t f_rows = fr;
        if (m->P > 1) {
            MB_STEP(reserve(c, m->S, m->S_cap, p, (size_t)m->P * cap * n));
            MB_STEP(all_gather(m, p, m->F[p], cap * n, m->S[p]));
            // the factor of the stack is replicated: identical input bits -> identical rank and factors on every GPU
            if (plain) { rc = pb_tsqr_r(c, m->S[p], m->P * cap, n, n, m->F[p]); f_rows = n; }
            else if (m->P * cap <= std::min<int64_t>(512, n)) { F = m->S[p]; f_rows = m->P * cap; rc = PB_OK; }   // a short stack is a valid factor as it is
            else rc = pb_tsqr_factor(c, m->S[p], m->P * cap, n, n, m->F[p], &f_rows);
            MB_STEP(rc);
        }
        MB_STEP(plain ? pb_pod_from_r(c, F, n, eps, n_L_in, &nl) : pb_pod_from_factor(c, F, f_rows, n, n, eps, n_L_in, &nl));
        c->wide = false;
        if (p == 0) { m->n_L = nl; m->keep = 0; }
        return PB_OK;
    }
    MB_STEP(pb_reserve(c, &c->Gbuf, &c->G_cap, (size_t)n * n));
    int rc;
    if (rows <= 0) rc = pb_memset(c, c->Gbuf, 0, (size_t)n * n * sizeof(double));
    else if (U_host) rc = pb_gram_host(c, U_host + row0 * ldu_host, rows, n, ldu_host, Udev, c->Gbuf);
    else rc = pb_gram(c, Udev, rows, n, ld, c->Gbuf);
    MB_STEP(rc);
    MB_STEP(all_reduce(m, p, c->Gbuf, n * n));
    int keep = 0;
    c->gram_rows = *std::max_element(m->rows.begin(), m->rows.end());   // the refinement decision must be the same on every rank
    MB_STEP(pb_pod_from_gram(c, c->Gbuf, n, eps, n_L_in, mode, &nl, &keep));
    c->wide = false;
    if (keep > 0) {
        rc = pb_reserve(c, &c->Ybuf, &c->Y_cap, (size_t)std::max<int64_t>(rows, 1) * keep);
        if (rc == PB_OK) rc = pb_reserve(c, &c->G2buf, &c->G2_cap, (size_t)keep * keep);
        if (rc == PB_OK) rc = rows > 0 ? pb_pod_pass2_gram(c, Udev, rows, ld, c->Ybuf, c->G2buf)
                                       : pb_memset(c, c->G2buf, 0, (size_t)keep * keep * sizeof(double));
        MB_STEP(rc);
        MB_STEP(all_reduce(m, p, c->G2buf, (int64_t)keep * keep));
        MB_STEP(pb_pod_pass2_finish(c, c->G2buf, eps, n_L_in, &nl));
    }
    if (p == 0) { m->n_L = nl; m->keep = keep; }
    return PB_OK;
}

// V_p (rows_p x n_L, ldv) of rank p's slab, with the gram2 refinement step when the decomposition asked for it
int basis_rank(pb_multi* m, int p, double* V, int64_t ldv) {
    pb_ctx* c = m->ctx[p];
    const int64_t rows = m->rows[p], n = m->n_st; const int nl = m->n_L;
    const double* Udev = m->U[p];
    int rc = PB_OK;
    if (rows > 0) rc = pb_pod_basis(c, Udev, rows, m->ldu[p], m->keep > 0 ? c->Ybuf : nullptr, V, ldv);
    MB_STEP(rc);
    int steps = 0;
    pb_pod_refine_steps(c, &steps);
    if (m->keep > 0 && steps > 0) {
        rc = pb_reserve(c, &c->Pbuf, &c->P_cap, (size_t)n * nl + (size_t)nl * nl);
        double* B = c->Pbuf; double* G3 = B ? B + (size_t)n * nl : nullptr;
        if (rc == PB_OK) rc = rows > 0 ? pb_project(c, V, ldv, nl, Udev, m->ldu[p], rows, n, B) : pb_memset(c, B, 0, (size_t)n * nl * sizeof(double));
        MB_STEP(rc);
        MB_STEP(all_reduce(m, p, B, n * nl));
        rc = rows > 0 ? pb_pod_refine_apply(c, Udev, rows, m->ldu[p], B, G3) : pb_memset(c, G3, 0, (size_t)nl * nl * sizeof(double));
        MB_STEP(rc);
        MB_STEP(all_reduce(m, p, G3, (int64_t)nl * nl));
        rc = rows > 0 ? pb_pod_refine_finish(c, G3, rows, V, ldv) : PB_OK;
        MB_STEP(rc);
    }
    return PB_OK;
}

int project_rank(pb_multi* m, int p, const double* V, int64_t ldv, int nl, const double* U, int64_t ldu, int64_t rows, int64_t n, double* v) {
    pb_ctx* c = m->ctx[p];
    int rc = rows > 0 ? pb_project(c, V, ldv, nl, U, ldu, rows, n, v) : pb_memset(c, v, 0, (size_t)n * nl * sizeof(double));
    MB_STEP(rc);
    MB_STEP(all_reduce(m, p, v, n * nl));
    return PB_OK;
}

int check_shape(pb_multi* m, int64_t n_h, int64_t n_st) {
    if (n_h <= 0 || n_st <= 0) { m->err = "pb_multi: bad shape"; return PB_ERR_SHAPE; }
    if (n_h < n_st) { m->err = "pb_multi: the row-sharded POD needs n_h >= n_st (use pb_pod on one GPU for wide matrices)"; return PB_ERR_SHAPE; }
    if (n_st > 8192) { m->err = "pb_multi: n_st exceeds 8192"; return PB_ERR_SHAPE; }
    return PB_OK;
}

}  // namespace

extern "C" {

int pb_multi_create(int n_gpus, const int* device_ids, pb_multi** out) {
    if (!out) return PB_ERR_ARG;
    *out = nullptr;
    if (n_gpus < 1 || n_gpus > MAX_GPUS) { pb_set_global_error("pb_multi_create: n_gpus outside [1, 16]"); return PB_ERR_ARG; }
    pb_multi* m = new pb_multi();
    m->P = n_gpus;
    for (int p = 0; p < n_gpus; ++p) m->dev.push_back(device_ids ? device_ids[p] : p);
    for (int p = 0; p < n_gpus; ++p) {
        pb_ctx* c = nullptr;
        int rc = pb_create(m->dev[p], &c);
        if (rc != PB_OK) { for (pb_ctx* x : m->ctx) pb_destroy(x); delete m; return rc; }
        m->ctx.push_back(c);
    }
    // peer access between every pair of distinct devices (logical ranks may share a device: tests on a one-GPU box)
    for (int p = 0; p < n_gpus; ++p) {
        cudaSetDevice(m->dev[p]);
        for (int q = 0; q < n_gpus; ++q) {
            if (m->dev[q] == m->dev[p]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->dev[p], m->dev[q]);
            if (!can) {
                pb_set_global_error("pb_multi_create: GPU " + std::to_string(m->dev[p]) + " cannot access GPU " + std::to_string(m->dev[q]) + " (no peer path)");
                for (pb_ctx* x : m->ctx) pb_destroy(x); delete m; return PB_ERR_CUDA;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                pb_set_global_error(std::string("pb_multi_create: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                for (pb_ctx* x : m->ctx) pb_destroy(x); delete m; return PB_ERR_CUDA;
            }
            cudaGetLastError();
        }
    }
    const int P = n_gpus;
    m->rc.assign(P, PB_OK); m->step_rc.assign(P, PB_OK); m->xptr.assign(P, nullptr); m->xval.assign(P, 0);
    m->U.assign(P, nullptr); m->rows.assign(P, 0); m->ldu.assign(P, 0);
    m->V.assign(P, nullptr); m->V_cap.assign(P, 0); m->ldv.assign(P, 0);
    m->F.assign(P, nullptr); m->F_cap.assign(P, 0); m->S.assign(P, nullptr); m->S_cap.assign(P, 0);
    m->vb.assign(P, nullptr); m->vb_cap.assign(P, 0);
    m->t0.assign(P, nullptr); m->t1.assign(P, nullptr);
    for (int k = 0; k < 3; ++k) m->ev[k].assign(P, nullptr);
    for (int p = 0; p < P; ++p) {
        cudaSetDevice(m->dev[p]);
        for (int k = 0; k < 3; ++k) cudaEventCreateWithFlags(&m->ev[k][p], cudaEventDisableTiming);
        cudaEventCreate(&m->t0[p]); cudaEventCreate(&m->t1[p]);
    }
    m->bar = new SpinBarrier(P);
    for (int p = 0; p < P; ++p) m->th.emplace_back(worker, m, p);
    *out = m;
    return PB_OK;
}

int pb_multi_destroy(pb_multi* m) {
    if (!m) return PB_OK;
    { std::lock_guard<std::mutex> g(m->mu); m->quit = true; }
    m->cv.notify_all();
    for (auto& t : m->th) t.join();
    for (int p = 0; p < m->P; ++p) {
        cudaSetDevice(m->dev[p]);
        cudaStreamSynchronize(m->ctx[p]->stream);
        for (double* b : {m->V[p], m->F[p], m->S[p], m->vb[p]}) if (b) cudaFree(b);
        for (int k = 0; k < 3; ++k) if (m->ev[k][p]) cudaEventDestroy(m->ev[k][p]);
        cudaEventDestroy(m->t0[p]); cudaEventDestroy(m->t1[p]);
        pb_destroy(m->ctx[p]);
    }
    delete m->bar;
    delete m;
    return PB_OK;
}

const char* pb_multi_last_error(const pb_multi* m) { return m ? m->err.c_str() : pb_last_error(nullptr); }
int pb_multi_size(const pb_multi* m) { return m ? m->P : 0; }
pb_ctx* pb_multi_ctx(pb_multi* m, int rank) { return (m && rank >= 0 && rank < m->P) ? m->ctx[rank] : nullptr; }

int pb_multi_row_range(const pb_multi* m, int64_t n_rows, int rank, int64_t* lo, int64_t* hi) {
    if (!m || !lo || !hi || rank < 0 || rank >= m->P) return PB_ERR_ARG;
    row_range(n_rows, rank, m->P, lo, hi);
    return PB_OK;
}

int pb_multi_sync(pb_multi* m) {
    if (!m) return PB_ERR_ARG;
    return run(m, [m](int p) { return pb_sync(m->ctx[p]); });
}

int pb_multi_all_reduce(pb_multi* m, double* const* bufs, int64_t count) {
    if (!m || !bufs || count <= 0) return PB_ERR_ARG;
    return run(m, [=](int p) { return all_reduce(m, p, bufs[p], count); });
}

int pb_multi_pod(pb_multi* m, const double* const* U_dev, const int64_t* rows, int64_t n_st, const int64_t* ldu,
                 double eps, int n_L_in, int mode, int* n_L) {
    if (!m || !U_dev || !rows || !n_L) return PB_ERR_ARG;
    int64_t n_h = 0;
    for (int p = 0; p < m->P; ++p) {
        if (rows[p] < 0 || (rows[p] > 0 && !U_dev[p]) || (ldu && rows[p] > 0 && ldu[p] < n_st)) { m->err = "pb_multi_pod: bad slab"; return PB_ERR_SHAPE; }
        n_h += rows[p];
    }
    PB_TRY(check_shape(m, n_h, n_st));
    for (int p = 0; p < m->P; ++p) { m->U[p] = U_dev[p]; m->rows[p] = rows[p]; m->ldu[p] = ldu ? ldu[p] : n_st; }
    m->n_h = n_h; m->n_st = n_st; m->mode = mode; m->have = false; m->host_slabs = false;
    PB_TRY(run(m, [=](int p) { return pod_rank(m, p, nullptr, 0, 0, eps, n_L_in, mode); }));
    m->have = true;
    *n_L = m->n_L;
    return PB_OK;
}

int pb_multi_pod_host(pb_multi* m, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host, double eps, int n_L_in,
                      int mode, int* n_L) {
    if (!m || !U_host || !n_L) return PB_ERR_ARG;
    if (ldu_host < n_st) { m->err = "pb_multi_pod_host: ldu_host < n_st"; return PB_ERR_SHAPE; }
    PB_TRY(check_shape(m, n_h, n_st));
    m->n_h = n_h; m->n_st = n_st; m->mode = mode; m->have = false; m->host_slabs = true;
    std::vector<int64_t> lo(m->P), hi(m->P);
    for (int p = 0; p < m->P; ++p) { row_range(n_h, p, m->P, &lo[p], &hi[p]); m->rows[p] = hi[p] - lo[p]; m->ldu[p] = n_st; }
    PB_TRY(run(m, [&, eps, n_L_in, mode](int p) {
        void* slab = nullptr;
        MB_STEP(pb_slab(m->ctx[p], (size_t)std::max<int64_t>(m->rows[p], 1) * n_st * sizeof(double), &slab));
        m->U[p] = static_cast<const double*>(slab);
        return pod_rank(m, p, U_host, ldu_host, lo[p], eps, n_L_in, mode);
    }));
    m->have = true;
    *n_L = m->n_L;
    return PB_OK;
}

// Two-step POD (pod.py:51-68), row-sharded: per-sample PODs with eps_init as ONE batch on every GPU's rows -- the batched
// n_t x n_t Gram partials are summed over the GPUs (one all-reduce per pass), the small eigensolves are replicated --
// then the global POD of the packed bases (n_h x sum r_k, tall) through the sharded path above.  ranks_host (nullable):
// the per-sample ranks.  pb_multi_basis afterwards gives every GPU's rows of V.
int pb_multi_fast_pod(pb_multi* m, const double* const* U_dev, const int64_t* rows, int n_t, int64_t n_s, const int64_t* ldu,
                      double eps, double eps_init, int mode, int* n_L, int* ranks_host) {
    if (!m || !U_dev || !rows || !n_L) return PB_ERR_ARG;
    if (n_t <= 0 || n_s <= 0 || n_t > 1024 || n_s >= 65536) { m->err = "pb_multi_fast_pod: bad shape (n_t in [1, 1024], n_s < 65536)"; return PB_ERR_SHAPE; }
    if (mode < PB_POD_GRAM || mode > PB_POD_TSQR) { m->err = "pb_multi_fast_pod: unknown mode"; return PB_ERR_ARG; }
    int64_t n_h = 0;
    for (int p = 0; p < m->P; ++p) {
        if (rows[p] < n_t) { m->err = "pb_multi_fast_pod: every GPU needs at least n_t rows"; return PB_ERR_SHAPE; }
        n_h += rows[p];
    }
    const int64_t n_st = (int64_t)n_t * n_s;
    m->have = false; m->host_slabs = false;
    std::vector<std::vector<int>> ranks(m->P, std::vector<int>((size_t)n_s, 0));
    std::vector<int64_t> total(m->P, 0);
    PB_TRY(run(m, [&, eps, eps_init, mode](int p) {
        pb_ctx* c = m->ctx[p];
        bool peer_abort = false;       // a peer failed before the collective: this rank has already taken part in that agreement
        std::function<int(double*, int64_t)> reduce = [m, p, &peer_abort](double* buf, int64_t count) {
            int rc = agree(m, p, PB_OK);
            if (rc != PB_OK) { peer_abort = true; return rc; }
            return all_reduce(m, p, buf, count);
        };
        {
            int rc = pbi_fast_pod_stage1(c, U_dev[p], rows[p], n_t, n_s, ldu ? ldu[p] : n_st, eps_init, mode, ranks[p].data(), &total[p], reduce);
            if (peer_abort) return rc;
            MB_STEP(rc);
        }
        // stage 2 on the packed bases (identical column count on every rank: the eigensolves are replicated)
        if (n_h < total[p]) { c->err = "pb_multi_fast_pod: the packed bases are wide (sum r_k > n_h): run the global stage on one GPU"; MB_STEP(PB_ERR_SHAPE); }
        if (total[p] > 8192) { c->err = "pb_multi_fast_pod: sum r_k exceeds 8192"; MB_STEP(PB_ERR_SHAPE); }
        m->U[p] = c->Uf; m->rows[p] = rows[p]; m->ldu[p] = n_st;
        if (p == 0) { m->n_h = n_h; m->n_st = total[0]; m->mode = mode; }
        MB_STEP(PB_OK);                                     // (publishes n_st before pod_rank reads it)
        return pod_rank(m, p, nullptr, 0, 0, eps, 0, mode);
    }));
    m->have = true;
    if (ranks_host) for (int64_t z = 0; z < n_s; ++z) ranks_host[z] = ranks[0][z];
    *n_L = m->n_L;
    return PB_OK;
}

int pb_multi_basis(pb_multi* m, double* const* V_dev, const int64_t* ldv) {
    if (!m || !V_dev) return PB_ERR_ARG;
    if (!m->have) { m->err = "pb_multi_basis: no decomposition (call pb_multi_pod first)"; return PB_ERR_ARG; }
    for (int p = 0; p < m->P; ++p)
        if ((m->rows[p] > 0 && !V_dev[p]) || (ldv && ldv[p] < m->n_L)) { m->err = "pb_multi_basis: bad V slab"; return PB_ERR_SHAPE; }
    return run(m, [=](int p) { return basis_rank(m, p, V_dev[p], ldv ? ldv[p] : m->n_L); });
}

int pb_multi_basis_host(pb_multi* m, double* V_host, int64_t ldv_host) {
    if (!m || !V_host) return PB_ERR_ARG;
    if (!m->have) { m->err = "pb_multi_basis_host: no decomposition (call pb_multi_pod_host first)"; return PB_ERR_ARG; }
    if (ldv_host < m->n_L) { m->err = "pb_multi_basis_host: ldv_host < n_L"; return PB_ERR_SHAPE; }
    const int nl = m->n_L;
    return run(m, [=](int p) {
        pb_ctx* c = m->ctx[p];
        MB_STEP(reserve(c, m->V, m->V_cap, p, (size_t)m->rows[p] * nl));
        m->ldv[p] = nl;
        PB_TRY(basis_rank(m, p, m->V[p], nl));
        int64_t lo, hi; row_range(m->n_h, p, m->P, &lo, &hi);
        if (!m->host_slabs) { lo = 0; for (int q = 0; q < p; ++q) lo += m->rows[q]; }
        if (m->rows[p] <= 0) return (int)PB_OK;
        if (ldv_host == nl) return pb_download(c, V_host + lo * ldv_host, m->V[p], (size_t)m->rows[p] * nl * sizeof(double));
        PB_CUDA(c, cudaMemcpy2DAsync(V_host + lo * ldv_host, ldv_host * sizeof(double), m->V[p], nl * sizeof(double), nl * sizeof(double),
                                     m->rows[p], cudaMemcpyDeviceToHost, c->stream));
        return pb_sync(c);
    });
}

int pb_multi_project(pb_multi* m, const double* const* V_dev, const int64_t* ldv, int n_L, const double* const* U_dev,
                     const int64_t* ldu, const int64_t* rows, int64_t n, double* const* v_dev) {
    if (!m || !V_dev || !U_dev || !rows || !v_dev || n_L <= 0 || n <= 0) return PB_ERR_ARG;
    return run(m, [=](int p) {
        return project_rank(m, p, V_dev[p], ldv ? ldv[p] : n_L, n_L, U_dev[p], ldu ? ldu[p] : n, rows[p], n, v_dev[p]);
    });
}

// v (n_st x n_L, host) = (V^T U)^T of the slabs and basis the last pb_multi_pod_host / pb_multi_basis_host left in HBM
int pb_multi_project_host(pb_multi* m, double* v_host) {
    if (!m || !v_host) return PB_ERR_ARG;
    if (!m->have || !m->V[0] || m->ldv[0] != m->n_L) { m->err = "pb_multi_project_host: call pb_multi_pod_host and pb_multi_basis_host first"; return PB_ERR_ARG; }
    const int nl = m->n_L; const int64_t n = m->n_st;
    return run(m, [=](int p) {
        pb_ctx* c = m->ctx[p];
        MB_STEP(reserve(c, m->vb, m->vb_cap, p, (size_t)n * nl));
        PB_TRY(project_rank(m, p, m->V[p], nl, nl, m->U[p], m->ldu[p], m->rows[p], n, m->vb[p]));
        if (p == 0) return pb_download(c, v_host, m->vb[0], (size_t)n * nl * sizeof(double));
        return pb_sync(c);
    });
}

int pb_multi_sigma(pb_multi* m, double* sigma_host, int64_t capacity, int64_t* count) {
    if (!m) return PB_ERR_ARG;
    return pb_pod_sigma(m->ctx[0], sigma_host, capacity, count);
}

// device time of the last pb_multi_pod / pb_multi_pod_host step is not kept; these bracket a region on every GPU:
// start records an event on every stream, stop records + synchronises and returns the MAX elapsed time over the GPUs
int pb_multi_timer_start(pb_multi* m) {
    if (!m) return PB_ERR_ARG;
    return run(m, [m](int p) { PB_CUDA(m->ctx[p], cudaEventRecord(m->t0[p], m->ctx[p]->stream)); return (int)PB_OK; });
}
int pb_multi_timer_stop(pb_multi* m, double* ms) {
    if (!m || !ms) return PB_ERR_ARG;
    std::vector<float> t(m->P, 0.f);
    PB_TRY(run(m, [&](int p) {
        pb_ctx* c = m->ctx[p];
        PB_CUDA(c, cudaEventRecord(m->t1[p], c->stream));
        PB_CUDA(c, cudaEventSynchronize(m->t1[p]));
        PB_CUDA(c, cudaEventElapsedTime(&t[p], m->t0[p], m->t1[p]));
        return (int)PB_OK;
    }));
    *ms = *std::max_element(t.begin(), t.end());
    return PB_OK;
}

}  // extern "C"
