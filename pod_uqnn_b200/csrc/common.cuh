// Shared declarations of libpodb200 (sm_100a).  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <functional>
#include "../../include/podb200.h"

constexpr int PB_STAGE_SLOTS = 3;

struct pb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    double ms[PB_T_COUNT] = {0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    cudaEvent_t evk0 = nullptr, evk1 = nullptr;     // around the TN main kernel
    cudaEvent_t evr0 = nullptr, evr1 = nullptr;     // user region timer
    cudaEvent_t tev[PB_T_COUNT][2] = {{nullptr}};   // stage timers: recorded on the stream, read lazily by pb_last_ms
    bool t_pending[PB_T_COUNT] = {false};           // (no host synchronisation inside the ops)
    bool pass2_t_started = false;        // pb_pod_pass2_gram recorded the start event of PB_T_PASS2
    bool sigma_pending = false;                     // h_sig (pinned) is being filled by an async copy; pb_pod_sigma finishes it
    double* h_sig = nullptr; size_t h_sig_cap = 0; int64_t sig_count = 0, sig_total = 0;
    bool tn_timing_pending = false;
    cudaStream_t copy_stream = nullptr;             // pb_pod_host: chunked upload overlapped with the Gram
    bool fork_pending = false;                      // pb_download_forked: a D2H copy is in flight on copy_stream (pb_sync joins it)
    cudaEvent_t chunk_ev[16] = {nullptr};
    double* Gpart = nullptr; size_t Gpart_cap = 0;
    void* stage[PB_STAGE_SLOTS] = {nullptr};        // pinned staging ring for pageable host inputs (api.cu)
    cudaEvent_t stage_ev[PB_STAGE_SLOTS] = {nullptr}; size_t stage_cap = 0;
    double* slab = nullptr; size_t slab_cap = 0;    // cached device slab of the host-facing calls (pb_slab)

    // split-K workspace of the TN products (grown on demand)
    double* ws = nullptr;  size_t ws_bytes = 0;
    // L2 flush scratch
    void* flush = nullptr; size_t flush_bytes = 0;

    // ---- state of the last decomposition (pb_pod_from_gram / pass2) ----
    int64_t m = 0;          // order of the Gram (columns of the tall operand)
    int mode = 0;
    int r = 0;              // numerical rank (vectors of the factor)
    int r2 = 0;             // rank of the second-pass factor
    int n_L = 0;
    int n_keep = 0;         // rotated columns for the second pass
    bool pass2_done = false;
    bool wide = false;      // pb_pod() transposed the input (n_h < n_st)
    int sweeps1 = 0, sweeps2 = 0;
    double* At = nullptr;   // (rows x ldm) factor vectors of pass 1 (unsorted; perm gives the order)
    double* At2 = nullptr;  // factor vectors of pass 2
    double* sig = nullptr;  // device (m): singular values, descending, zero beyond r
    double* nrm = nullptr;  // device (m): norms of the factor vectors in the same order (incl. shift)
    double* sig2 = nullptr; double* nrm2 = nullptr;
    int* perm = nullptr;    // device (m + 32): vector index of the k-th largest
    int* perm2 = nullptr;
    double* Bmat = nullptr; // right-factor scratch
    double* Wk = nullptr;   // W_keep (m x n_keep) row-major
    double* small = nullptr;// eigensolver scratch
    double* Gbuf = nullptr; double* G2buf = nullptr; double* Ybuf = nullptr; double* ATbuf = nullptr;
    double* Uf = nullptr; size_t Uf_cap = 0; int64_t uf_rows = 0, uf_cols = 0, uf_ld = 0;  // two-step POD stage-1 bases
    double* fp_buf[8] = {nullptr}; size_t fp_cap[8] = {0};   // batched two-step POD work buffers
    int* fp_ibuf = nullptr; size_t fp_icap = 0;
    size_t At_cap = 0, At2_cap = 0, sig_cap = 0, nrm_cap = 0, sig2_cap = 0, nrm2_cap = 0, B_cap = 0,
           Wk_cap = 0, small_cap = 0, G_cap = 0, G2_cap = 0, Y_cap = 0, AT_cap = 0;
    size_t perm_cap = 0, perm2_cap = 0;
    int* d_int = nullptr;   // device ints (flags, n_L ...)
    int* h_int = nullptr;   // pinned host mirror
    std::vector<double> sigma_host;  // singular values of the last decomposition
    double* qr_W = nullptr; double* qr_scr = nullptr; double* qr_scr2 = nullptr;   // TSQR mode (qr.cu)
    double* qr_stack = nullptr; size_t qr_stack_cap = 0;
    int64_t qr_panels = 0; double qr_flops = 0.; int64_t qr_f_rows = 0;   // last factorisation: panels run, Householder flops, factor rows
    size_t qr_W_cap = 0, qr_scr_cap = 0, qr_scr2_cap = 0;
    // per-call knobs of the next eigensolve (set by the POD orchestration, consumed and cleared by eig.cu)
    bool eig_trust_single_pair = false;   // skip the confirming sweep of a single-pair problem (a second pass follows)
    double eig_jacobi_tol = 0.;           // > 0: rotation threshold of the next eig (first pass of the accurate mode)
    int refine_steps = 0; double refine_est = 0.; int64_t gram_rows = 0; int64_t refine_rows_fixed = 0;   // gram2: subspace-refinement decision of the last decomposition (api.cu)
    double* Rbuf = nullptr; size_t R_cap = 0;       // refinement: V~ (rows x n_L)
    double* Pbuf = nullptr; size_t P_cap = 0;       // refinement: projection block + small Gram
    void* gram_plan = nullptr;            // stream-K schedule of the last Gram shape (gram_tma.cu), owned by the context
    void* train_state = nullptr;     // pb_ensemble_train: work buffers, side stream, captured epoch graph (train.cu)
};
void pbi_train_free(pb_ctx* ctx);
void pbi_gram_plan_free(pb_ctx* ctx);

const char* pb_set_global_error(const std::string& s);

#define PB_FAIL(ctx, code, ...) do { char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__); \
    if (ctx) (ctx)->err = _b; else pb_set_global_error(_b); return (code); } while (0)

#define PB_CUDA(ctx, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
    PB_FAIL(ctx, PB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); } } while (0)

// every extern "C" entry point that takes a context: null check + make the context's GPU the calling thread's current
// device (the caller may be a new Python thread, or torch may have switched devices since the last call)
#define PB_ENTER(ctx) do { if (!(ctx)) return PB_ERR_ARG; int _dev = -1; \
    if (cudaGetDevice(&_dev) != cudaSuccess || _dev != (ctx)->device) PB_CUDA(ctx, cudaSetDevice((ctx)->device)); } while (0)

#define PB_TRY(expr) do { int _r = (expr); if (_r != PB_OK) return _r; } while (0)

#define PB_LAUNCH_CHECK(ctx) do { (ctx)->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { \
    PB_FAIL(ctx, PB_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); } } while (0)

// Opt `func` in to the device's maximal dynamic shared memory (once per function and device, cached) and check that
// `need` fits.  The limit is raised to the maximum, never to `need`: an exact size set per launch is a race when two
// contexts drive the same device from two threads (one lowers the limit between the other's attribute call and launch).
int pbi_smem_optin(pb_ctx* ctx, const void* func, size_t need);

// grow-on-demand device buffer
int pb_reserve(pb_ctx* ctx, double** p, size_t* cap, size_t elems);

// ---- internal kernels' host entry points (one per .cu) ----
// C (ka x kb, ldc) = A^T B over rows; sym != 0: A == B, only tiles j >= i computed then mirrored.
int pbi_gemm_tn(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka,
                const double* B, int64_t ldb, int64_t kb, int64_t rows,
                double* C, int64_t ldc, int sym);
// Y (rows x k, ldy) = A (rows x m, lda) B (m x k, ldb); optional fused pod_sig epilogue:
// if Uref != nullptr: sig_out[row] = mean_j |Uref - Y| / 2 (Y itself stored only when Y != nullptr)
int pbi_gemm_nn(pb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t m,
                const double* B, int64_t ldb, int64_t k, double* Y, int64_t ldy,
                const double* Uref, int64_t ldu, double* sig_out);
// small dense helpers
int pbi_transpose(pb_ctx* ctx, const double* in, int64_t rows, int64_t cols, int64_t ldin,
                  double* out, int64_t ldout);
// eigen-decomposition of a PSD matrix G (m x m, ld m): factor vectors At (r x m) sorted by
// descending norm, sig[0..m) (zero beyond r).  tol_rel: Cholesky truncation (relative to max diag),
// shift_rel: diagonal shift relative to trace (full-rank mode), returns r.
int pbi_eig_psd_small_queued(pb_ctx* ctx, const double* G, int64_t m, double** At_buf, size_t* At_cap, double* sig_dev,
                             double* nrm_dev, int* perm_dev, double eps);
int pbi_eig_psd(pb_ctx* ctx, const double* G, int64_t m, double tol_rel, double shift_rel,
                double** At_buf, size_t* At_cap, double* sig_dev, double* nrm_dev, int* perm_dev,
                int* r_out, int* sweeps_out, int max_sweeps);
// one-sided block Jacobi + norms + sort on r factor vectors already in At (shift_dev == nullptr: no shift)
int pbi_jacobi_factor(pb_ctx* ctx, double* At, int64_t m, int r, const double* shift_dev, double* sig_dev, double* nrm_dev,
                      int* perm_dev, int* sweeps_out, int max_sweeps);
// TSQR mode (qr.cu): R factor of a tall matrix; rank-revealing pivoted QR of the small factor held in At
int pbi_qr_r(pb_ctx* ctx, const double* A, int64_t rows, int64_t m, int64_t lda, double* R, int64_t ldr, int64_t rows_out);
// rr: rank-revealing (block pivoting + early termination): *f_rows <= m dense rows in the input's column order
int pbi_qr_factor(pb_ctx* ctx, const double* A, int64_t rows, int64_t m, int64_t lda, double* R, int64_t ldr, int64_t rows_out,
                  bool rr, int64_t* f_rows);
int pbi_qr_load_factor(pb_ctx* ctx, const double* R, int64_t ldr, int64_t m, double* At, int64_t ldm, int64_t rows_total);
int pbi_qr_load_dense(pb_ctx* ctx, const double* F, int64_t ldf, int64_t f_rows, int64_t m, double* At, int64_t ldm, int64_t rows_total);
int pbi_qrcp_factor(pb_ctx* ctx, double* At, int64_t m, int64_t nrows, int64_t ldm, int64_t rows_total, double tol_rel, int* r_out);
// Y (rows x k, ldy) += A (rows x m, lda) B (m x k, ldb), in place
int pbi_gemm_nn_sub(pb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t m, const double* B, int64_t ldb,
                    int64_t k, double* Y, int64_t ldy);
// n_L from descending lambdas = sig^2 (pod.py:22-32)
int pbi_rank(pb_ctx* ctx, const double* sig_dev, int64_t count, double eps, int* n_L_out);
// B (m x k, ldb) [i][j] = At[perm[j]][i] / (nrm[j] * (sig ? sig[j] : 1))
int pbi_factor_to_right(pb_ctx* ctx, const double* At, int64_t m, int k, const int* perm_dev,
                        const double* nrm_dev, const double* sig_dev, double* B, int64_t ldb);
int pb_reserve_int(pb_ctx* ctx, int** p, size_t* cap, size_t elems);
// C (rows x n, ldc) += Y (rows x 64, ldy) B (64 x n, ldb) in place; 1 = outside the kernel's domain (qr_update.cu)
int pbi_qr_update(pb_ctx* ctx, const double* Y, int64_t ldy, int64_t rows, int64_t K, const double* B, int64_t ldb, int64_t n,
                  double* C, int64_t ldc);
int pbi_fast_pod_stage1(pb_ctx* ctx, const double* U, int64_t n_h_loc, int n_t, int64_t n_s, int64_t ldu, double eps_init,
                        int mode, int* ranks_host, int64_t* total, const std::function<int(double*, int64_t)>& reduce);
// batched variants (two-step POD)
int pbi_gram_batched(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka, int64_t sA, int64_t rows,
                     double* C, int64_t sC, int count);
int pbi_gemm_nn_batched(pb_ctx* ctx, const double* A, int64_t lda, int64_t sA, int64_t rows, int64_t m,
                        const double* B, int64_t ldb, int64_t sB, int64_t k, double* Y, int64_t ldy, int64_t sY, int count);
int pbi_eig_psd_batched(pb_ctx* ctx, const double* G_all, int64_t m, int batch, double tol_rel, double shift_rel,
                        double* At_all, double* sig_all, double* nrm_all, int* perm_all, int* r_dev,
                        double* scratch_dbl, int* sweeps_out, int max_sweeps);
int pbi_rank_batched(pb_ctx* ctx, const double* sig_all, int64_t m, int batch, double eps, int* nL_dev);
int pbi_factor_to_right_batched(pb_ctx* ctx, const double* At_all, int64_t m, int k, int batch, const int* perm_all,
                                const double* nrm_all, const double* sig_all, double* B_all, int64_t bsB);
