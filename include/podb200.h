/*
 * podb200.h -- C-ABI of libpodb200.so: the B200 (sm_100a) hot path of POD-UQNN.
 *
 * The reference (pierremtb/POD-UQNN) has no FFI; its seam is a set of Python
 * signatures.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference tree).  The Python shims in
 * pod_uqnn_b200/ bind these with ctypes and keep the reference's names and
 * argument meaning; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - every function returns int: PB_OK (0) or a negative PB_ERR_*;
 *     pb_last_error(ctx) gives the message of the last failure on that context.
 *   - all matrices are float64, row-major, with an explicit leading dimension
 *     (ld, in elements).  "dev" pointers are device pointers on the context's
 *     GPU (from pb_malloc or borrowed, e.g. torch.Tensor.data_ptr()); "host"
 *     pointers are ordinary host memory.
 *   - one context = one GPU, one stream.  Not thread-safe per context.
 *   - outputs are caller-allocated.  Where a size is only known after the call
 *     (the truncation rank n_L) the call is two-phase: pb_pod() returns n_L and
 *     keeps the factors in the context, pb_pod_basis() fills the caller buffer.
 *   - there is no CPU fallback: without a CUDA device pb_create() fails.
 */
#ifndef PODB200_H
#define PODB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_OK                      0
#define PB_ERR_CUDA               -1
#define PB_ERR_SHAPE              -2
#define PB_ERR_UNSUPPORTED_FIELD  -3
#define PB_ERR_NOT_CONVERGED      -4
#define PB_ERR_ARG                -5
#define PB_ERR_NOMEM              -6

/* analytic fields u(X, t, mu) (reference experiments/<case>/hyperparams.py) */
#define PB_FIELD_SHEKEL   1   /* 1d_shekel/hyperparams.py:55-67  */
#define PB_FIELD_ACKLEY   2   /* 2d_ackley/hyperparams.py:51-57  */
#define PB_FIELD_BURGERS  3   /* 1dt_burger/hyperparams.py:45-55 */

/* POD modes */
#define PB_POD_GRAM      0  /* method of snapshots, one Gram pass                     */
#define PB_POD_GRAM2     1  /* + second Gram pass on the rotated numerical-rank block */
#define PB_POD_ACCURATE  2  /* full rotate + second Gram (Jacobi-SVD accuracy class)  */
#define PB_POD_TSQR      3  /* Householder QR of U (never U^T U), pivoted QR + Jacobi SVD of the factor; rank-revealing:
                               block column pivoting, stops when the residual is below the QR's rounding error */
#define PB_POD_TSQR_FULL 4  /* the same without the early stop: all n_cols/64 panels (pb_pod only; benchmarking) */

/* input normalisation of the surrogate (varneuralnetwork.py:10-12,75-86) */
#define PB_NORM_NONE     0
#define PB_NORM_MEANSTD  1
#define PB_NORM_CENTER   2

/* timers readable with pb_last_ms() */
#define PB_T_GRAM       0
#define PB_T_EIG        1
#define PB_T_BASIS      2
#define PB_T_PROJECT    3
#define PB_T_RECON      4
#define PB_T_SNAPSHOTS  5
#define PB_T_PREDICT    6
#define PB_T_POD_TOTAL  7
#define PB_T_PASS2      8
#define PB_T_TN_KERNEL  9   /* the split-K DMMA kernel of the last pb_gram / TN product alone */
#define PB_T_QR         10  /* the blocked Householder QR of the last pb_tsqr_r / TSQR-mode pb_pod */
#define PB_T_REFINE     11  /* gram2: the subspace-refinement step of the last pb_pod_basis_full (0 when not needed) */
#define PB_T_COUNT      12

typedef struct pb_ctx pb_ctx;

/* ---- lifetime ---------------------------------------------------------- */
int  pb_create(int device, pb_ctx** out);
int  pb_destroy(pb_ctx* ctx);
const char* pb_last_error(const pb_ctx* ctx);       /* ctx may be NULL: global message */
int  pb_set_stream(pb_ctx* ctx, void* cuda_stream); /* NULL = the context's own stream */
int  pb_sync(pb_ctx* ctx);
int  pb_device_info(pb_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* mem_bytes);
int  pb_last_ms(pb_ctx* ctx, int which, double* ms);  /* device time of the last op of that kind */
int64_t pb_kernel_launches(pb_ctx* ctx);              /* kernels launched by this context so far */
/* region timer on the context stream (CUDA events): start, then stop returns the device ms */
int  pb_timer_start(pb_ctx* ctx);
int  pb_timer_stop(pb_ctx* ctx, double* ms);

/* ---- device memory (plain cudaMalloc / copies on the context stream) ---- */
int  pb_malloc(pb_ctx* ctx, size_t bytes, void** dev);
int  pb_free(pb_ctx* ctx, void* dev);
int  pb_memset(pb_ctx* ctx, void* dev, int value, size_t bytes);
int  pb_upload(pb_ctx* ctx, void* dev, const void* host, size_t bytes);   /* H2D + sync */
int  pb_download(pb_ctx* ctx, void* host, const void* dev, size_t bytes); /* D2H + sync */
int  pb_upload_async(pb_ctx* ctx, void* dev, const void* host, size_t bytes);
int  pb_download_async(pb_ctx* ctx, void* host, const void* dev, size_t bytes);
/* D2H on the context's copy stream: ordered after everything queued so far, concurrent with what is queued next;
 * pb_sync joins it.  (V travelling to the host while pb_project runs.) */
int  pb_download_forked(pb_ctx* ctx, void* host, const void* dev, size_t bytes);
/* Context-owned device slab for the host-facing calls (pb_pod_host / pb_gram_host receive the uploaded matrix in
 * it): grown on demand, reused across calls (a cudaMalloc / cudaFree pair of 1.28 GB per perform_pod call costs
 * milliseconds and synchronises the device).  Valid until a later pb_slab asks for more, or pb_destroy. */
int  pb_slab(pb_ctx* ctx, size_t bytes, void** dev);
/* Host matrix (n_h x n_st doubles, row pitch ldu_host) -> U_dev (row pitch n_st), queued on the context stream.  Pinned
 * sources are copied directly, pageable ones (numpy arrays) through the pinned staging ring filled by host threads. */
int  pb_upload_matrix(pb_ctx* ctx, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host, double* U_dev);
int  pb_host_alloc(size_t bytes, void** host);        /* pinned host memory */
int  pb_host_free(void* host);

/* ---- snapshot evaluation ------------------------------------------------
 * replaces poduqnn/acceleration.py:13-30 (loop_u) and :34-70 (loop_u_t) for
 * the registered fields, noise-free branch.
 *   X_dev   (dim, n_xyz) node coordinates = x_mesh[:, 1:].T
 *   mu_dev  (n_s, n_mu)  parameter samples
 *   t_dev   (n_t) time grid, or NULL / n_t = 0 for steady fields
 *   U_dev   (n_v*n_xyz, ldu >= n_s*max(n_t,1)) snapshot matrix; column of
 *           sample i, step j is i*n_t + j; row is v*n_xyz + node
 *   row0, n_rows: the node range [row0, row0+n_rows) this call fills (row
 *           sharding across GPUs); U_dev points at the first LOCAL row.
 */
int  pb_snapshots(pb_ctx* ctx, int field, int dim, int64_t n_xyz,
                  const double* X_dev, int64_t n_s, int n_mu, const double* mu_dev,
                  int n_t, const double* t_dev,
                  int64_t row0, int64_t n_rows, double* U_dev, int64_t ldu);

/* Ackley field with de-duplicated coordinates (tensor grids: n_x + n_y distinct values for n_x n_y nodes):
 * ux/uy = the distinct x / y values, ix/iy (int32, n_xyz) = index of every node's coordinate in them.
 * exp(.5 cos(w x)) and exp(.5 cos(w y)) are tabulated per (coordinate, sample); one exp per element is left
 * and the kernel runs at the HBM write roofline.  Same result as pb_snapshots to ~1e-16 of max|U|. */
int  pb_snapshots_ackley_indexed(pb_ctx* ctx, int64_t n_xyz, const double* X_dev, int n_ux, const double* ux_dev,
                                 const int* ix_dev, int n_uy, const double* uy_dev, const int* iy_dev,
                                 int64_t n_s, const double* mu_dev, int64_t row0, int64_t n_rows,
                                 double* U_dev, int64_t ldu);

/* Ackley field on radius-sorted nodes (symmetric meshes repeat r = sqrt(.5 (x^2 + y^2)); the first term
 * -20 (1 + .1 mu_2) exp(-.2 (1 + .1 mu_1) r) of experiments/2d_ackley/hyperparams.py:51-57 only depends on (r, sample)):
 * perm (int32, n_rows) = the node ids of [row0, row0 + n_rows) ordered by r, r_sorted (n_rows) = their radii.
 * One exp per run of equal radii instead of one per element; bit-identical to pb_snapshots_ackley_indexed. */
int  pb_snapshots_ackley_grouped(pb_ctx* ctx, int64_t n_xyz, int n_ux, const double* ux_dev, const int* ix_dev,
                                 int n_uy, const double* uy_dev, const int* iy_dev, int64_t n_s,
                                 const double* mu_dev, int64_t row0, int64_t n_rows, const int* perm_dev,
                                 const double* r_sorted_dev, double* U_dev, int64_t ldu);

/* Known-answer matrix U = sum_k s_k phi_k psi_k^T (DCT-II vectors, s_k = 2^(-decay k)),
 * rows [row0, row0+n_rows) of an n_h x n_s matrix, filled in HBM. */
int  pb_dct_lowrank(pb_ctx* ctx, int64_t n_h, int64_t n_s, int r, double decay,
                    int64_t row0, int64_t n_rows, double* U_dev, int64_t ldu);

/* ---- POD ----------------------------------------------------------------
 * replaces poduqnn/pod.py:7-47 (perform_pod).
 *
 * Single-GPU:   pb_pod(ctx, U, n_h, n_st, ld, eps, n_L_in, mode, &n_L)
 * Row-sharded:  every rank calls pb_gram() on its row slab, all-reduces the
 *               n_st x n_st Gram (NCCL sum), calls pb_pod_from_gram() (the
 *               small eigensolve is replicated and deterministic), and
 *               pb_pod_basis() on its slab.  For the second-pass modes the
 *               rank repeats with pb_pod_pass2_gram() / pb_pod_pass2_finish().
 *
 * pb_gram:  G (n_cols x n_cols, ld = n_cols) = A^T A over the n_rows local rows.
 */
int  pb_gram(pb_ctx* ctx, const double* A_dev, int64_t n_rows, int64_t n_cols, int64_t lda,
             double* G_dev);
/* C (ka x kb, ld = kb) = A^T B over n_rows rows (projection-shaped product). */
int  pb_gemm_tn(pb_ctx* ctx, const double* A_dev, int64_t lda, int64_t ka,
                const double* B_dev, int64_t ldb, int64_t kb, int64_t n_rows, double* C_dev);
/* Y (n_rows x k, ldy) = A (n_rows x m, lda) * B (m x k, ldb). */
int  pb_gemm_nn(pb_ctx* ctx, const double* A_dev, int64_t lda, int64_t n_rows, int64_t m,
                const double* B_dev, int64_t ldb, int64_t k, double* Y_dev, int64_t ldy);

/* Eigen-decomposition of the (all-reduced) Gram + truncation rank (pod.py:22-32).
 * mode as above; eps / n_L_in as perform_pod(U, eps, n_L).  On return *n_keep is the
 * number of rotated columns the second pass needs (0 for PB_POD_GRAM) and *n_L the
 * rank -- provisional (the numerical rank) when *n_keep > 0: the second pass
 * (pb_pod_pass2_gram / pb_pod_pass2_finish) decides sigma and n_L from the rotated
 * block, so the first pass spends no rank kernel / host round trip on them. */
int  pb_pod_from_gram(pb_ctx* ctx, const double* G_dev, int64_t n_cols, double eps, int n_L_in,
                      int mode, int* n_L, int* n_keep);
/* Second pass (modes GRAM2 / ACCURATE): Y = A W_keep on the local slab, G2 = Y^T Y.
 * Y_dev is caller-allocated (n_rows x n_keep, ld = n_keep), G2_dev (n_keep x n_keep). */
int  pb_pod_pass2_gram(pb_ctx* ctx, const double* A_dev, int64_t n_rows, int64_t lda,
                       double* Y_dev, double* G2_dev);
int  pb_pod_pass2_finish(pb_ctx* ctx, const double* G2_dev, double eps, int n_L_in, int* n_L);
/* V (n_rows x n_L, ldv) for the local slab.  A_dev is the slab of U (tall case).
 * After a second pass Y_dev must be the buffer filled by pb_pod_pass2_gram (else NULL). */
int  pb_pod_basis(pb_ctx* ctx, const double* A_dev, int64_t n_rows, int64_t lda,
                  const double* Y_dev, double* V_dev, int64_t ldv);
/* gram2 mode, subspace refinement (tolerance safety; see api.cu): the modes of a Gram-based POD carry an error along
 * directions the numerical-rank block of the first pass does not span.  pb_pod_refine_steps returns 1 when the
 * a-priori estimate 10 eps (sigma_1/sigma_nL)^3 sqrt(4 eps) exceeds north_star's 1e-8 angle tolerance; then ONE
 * step of subspace iteration with the matrix itself follows the basis: B = A^T V (pb_project on every slab +
 * all-reduce), pb_pod_refine_apply (V~ = A_local B Sigma^-2, G3 = V~^T V~ partial), all-reduce G3,
 * pb_pod_refine_finish (V = V~ R^-1, R^T R = G3).  pb_pod_basis_full does all of it on one GPU. */
int  pb_pod_refine_steps(pb_ctx* ctx, int* steps);
/* The row count the a-priori estimate of the decomposition in progress uses instead of the rows of its Gram (0 = those):
 * row-sharded callers pass the LARGEST slab of the job between pb_gram / pb_gram_host and pb_pod_from_gram, so that every
 * rank takes the same decision from replicated quantities only and no collective is needed for it.  Every new Gram
 * clears the override. */
int  pb_pod_set_refine_rows(pb_ctx* ctx, int64_t rows);
int  pb_pod_refine_estimate(pb_ctx* ctx, double* est);   /* the a-priori angle estimate the decision was taken on */
int  pb_pod_refine_apply(pb_ctx* ctx, const double* A_dev, int64_t n_rows, int64_t lda, double* B_dev, double* G3_dev);
int  pb_pod_refine_finish(pb_ctx* ctx, const double* G3_dev, int64_t n_rows, double* V_dev, int64_t ldv);
/* singular values of the last decomposition (descending; zero beyond numerical rank) */
int  pb_pod_sigma(pb_ctx* ctx, double* sigma_host, int64_t capacity, int64_t* count);
/* right factor of the last decomposition: Z (n_cols x n_L, ld = n_L), V = A Z diag(1/sigma) */
int  pb_pod_right(pb_ctx* ctx, double* Z_dev, int64_t ldz);

/* TSQR accuracy mode (north_star: "a TSQR accuracy mode for ill-conditioned snapshot sets"; the reference's SVD,
 * pod.py:16, works on U itself -- so does this mode, the Gram modes work on U^T U).
 *   pb_tsqr_r:     R (n_cols x n_cols, ld = n_cols, upper triangular, zeros below) of the local slab A
 *                  (n_rows x n_cols): blocked Householder QR, Q is not formed.  Row-sharded: every rank calls it on
 *                  its slab, all-gathers the P factors, stacks them (P n_cols x n_cols) and calls pb_tsqr_r again
 *                  on the stack (replicated) -- TSQR with one leaf per GPU.
 *   pb_pod_from_r: rank-revealing pivoted QR of R, one-sided Jacobi SVD of the factor, truncation rank
 *                  (pod.py:22-32).  Afterwards pb_pod_basis(A, ..., Y = NULL, V) gives V = A Z / sigma (pod.py:42-45)
 *                  and pb_pod_sigma / pb_pod_right work as after pb_pod_from_gram. */
int  pb_tsqr_r(pb_ctx* ctx, const double* A_dev, int64_t n_rows, int64_t n_cols, int64_t lda, double* R_dev);
int  pb_pod_from_r(pb_ctx* ctx, const double* R_dev, int64_t n_cols, double eps, int n_L_in, int* n_L);
/* Rank-revealing form (what PB_POD_TSQR runs).  Snapshot matrices have low numerical rank, and the SVD (pod.py:16)
 * needs only a factor F with F^T F = A^T A to working accuracy:
 *   pb_tsqr_factor:     blocked Householder QR with block column pivoting (before each 64-column panel the columns
 *                       with the largest residual norms are moved into it) that STOPS as soon as every residual column
 *                       is below 16 eps sqrt(panels) x the largest column norm of A -- the level of the rounding error
 *                       the reflections themselves leave in the kept rows.  F_dev is an n_cols x n_cols buffer (ld =
 *                       n_cols): rows [0, *f_rows) hold the factor (dense, in A's column order), the rest is zeroed.
 *                       A matrix of full numerical rank runs every panel and gets the triangular R, column-permuted.
 *                       Row-sharded: every rank factors its slab, the first max_p f_rows_p rows of the P factors are
 *                       all-gathered and stacked, pb_tsqr_factor again on the stack (replicated).
 *   pb_pod_from_factor: pivoted QR + Jacobi SVD of a dense factor (f_rows x n_cols, ld ldf), as pb_pod_from_r.
 *   pb_tsqr_stats:      of the last pb_tsqr_r / pb_tsqr_factor: 64-column panels run (all leaves), rows of the factor,
 *                       Householder flop count of the panels actually run. */
int  pb_tsqr_factor(pb_ctx* ctx, const double* A_dev, int64_t n_rows, int64_t n_cols, int64_t lda, double* F_dev,
                    int64_t* f_rows);
int  pb_pod_from_factor(pb_ctx* ctx, const double* F_dev, int64_t f_rows, int64_t n_cols, int64_t ldf, double eps,
                        int n_L_in, int* n_L);
int  pb_tsqr_stats(pb_ctx* ctx, int64_t* panels, int64_t* f_rows, double* flops);

/* One-call single-GPU POD of a device matrix (tall or wide).  V via pb_pod_basis_full(). */
int  pb_pod(pb_ctx* ctx, const double* U_dev, int64_t n_h, int64_t n_st, int64_t ldu,
            double eps, int n_L_in, int mode, int* n_L);
int  pb_pod_basis_full(pb_ctx* ctx, const double* U_dev, int64_t n_h, int64_t n_st, int64_t ldu,
                       double* V_dev, int64_t ldv);
/* Gram of a HOST matrix, for row-sharded callers (the all-reduce of G follows): chunked upload on a copy
 * stream overlapped with the per-chunk Gram, as pb_pod_host.  U_dev (n_h x n_st, ld n_st) receives the slab,
 * G_dev (n_st x n_st) its Gram.  (pod.py:16 applied to one rank's rows.) */
int  pb_gram_host(pb_ctx* ctx, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host,
                  double* U_dev, double* G_dev);
/* Same from a HOST matrix: the upload into U_dev (n_h x n_st, ld = n_st, caller-allocated) is cut into row chunks on
 * a copy stream and overlapped with the Gram.  Pinned memory (pb_host_alloc) is copied directly; pageable memory
 * (an ordinary numpy array, what poduqnn/pod.py:7 is handed) goes through a ring of pinned staging buffers filled by
 * host threads (PB_HOST_COPY_THREADS, default half the cores, at most 16).
 * Mode PB_POD_TSQR (rank-revealing): the leaf QR of row chunk c runs while chunks c + 1 .. cross PCIe (an uploader
 * thread feeds the copy stream, the calling thread drives the leaf factorisations and their pivot decisions), then the
 * QR of the stacked leaf factors; PB_POD_TSQR_FULL uploads first. */
int  pb_pod_host(pb_ctx* ctx, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host, double* U_dev,
                 double eps, int n_L_in, int mode, int* n_L);

/* replaces poduqnn/pod.py:51-68 (perform_fast_pod): per-sample POD with eps_init on the
 * n_h x n_t column blocks of U (sample-major columns, as loop_u_t writes them), unweighted
 * concatenation, global POD with eps.  ranks_host (nullable, n_s ints) gets the per-sample ranks. */
int  pb_fast_pod(pb_ctx* ctx, const double* U_dev, int64_t n_h, int n_t, int64_t n_s, int64_t ldu,
                 double eps, double eps_init, int mode, int* n_L, int* ranks_host);
int  pb_fast_pod_basis(pb_ctx* ctx, double* V_dev, int64_t ldv);   /* V (n_h x n_L) of the last pb_fast_pod */
/* (n_h, n_t, n_s) C-order U_struct  ->  (n_h, n_s*n_t) sample-major matrix (and back) */
int  pb_struct_to_matrix(pb_ctx* ctx, const double* Us_dev, int64_t n_h, int n_t, int64_t n_s,
                         double* U_dev, int64_t ldu);
int  pb_matrix_to_struct(pb_ctx* ctx, const double* U_dev, int64_t ldu, int64_t n_h, int n_t,
                         int64_t n_s, double* Us_dev);

/* ---- single-process multi-GPU (csrc/multi.cu) ------------------------------
 * "pb_create(n_gpus, device_ids) ... no MPI, no torchrun needed".  A pb_multi owns one context (stream +
 * work buffers) and one host thread per GPU.  The snapshot matrix is row-sharded by dof (contiguous, balanced slabs:
 * pb_multi_row_range); the exchange steps -- sum of the Gram / second-pass Gram / refinement / projection partials,
 * gather of the TSQR factors -- run as the library's own peer-memory kernels and copies over NVLink (deterministic:
 * every GPU ends with the same bits), the small eigensolve is replicated.  Stands behind the reference's
 * single-process calls pod.perform_pod(U, eps, n_L) (pod.py:7-47) and (V.T.dot(U)).T (podnnmodel.py:435-437) when
 * n_gpus > 1.  device_ids == NULL: GPUs 0..n_gpus-1.  Logical ranks may share a device (tests on a one-GPU box).
 *   pb_multi_ctx:          the context of rank p (for pb_malloc / pb_upload / pb_snapshots ... on that GPU)
 *   pb_multi_pod:          slabs already in HBM: U_dev[p] is rows[p] x n_st with pitch ldu[p] (NULL: n_st)
 *   pb_multi_pod_host:     host matrix in (pinned or pageable), slabs uploaded concurrently (chunked, overlapped with
 *                          the per-chunk Grams) and kept in each context's slab for the calls below
 *   pb_multi_basis:        V_dev[p] (rows[p] x n_L, pitch ldv[p] or n_L) = rank p's rows of the modes
 *   pb_multi_basis_host:   the same gathered into a host matrix (n_h x n_L, pitch ldv_host)
 *   pb_multi_project:      v_dev[p] (n x n_L) = sum over ranks of (V_p^T U_p)^T, replicated on every GPU
 *   pb_multi_project_host: projection of the matrix the last pb_multi_pod_host uploaded onto the basis the last
 *                          pb_multi_basis_host formed (both still in HBM) -> v_host (n_st x n_L)
 *   pb_multi_all_reduce:   in-place sum of bufs[p] (count doubles on GPU p) over the ranks
 *   pb_multi_timer_*:      bracket a region on every GPU's stream; stop returns the MAX device time over the GPUs */
typedef struct pb_multi pb_multi;
int  pb_multi_create(int n_gpus, const int* device_ids, pb_multi** out);
int  pb_multi_destroy(pb_multi* m);
const char* pb_multi_last_error(const pb_multi* m);
int  pb_multi_size(const pb_multi* m);
pb_ctx* pb_multi_ctx(pb_multi* m, int rank);
int  pb_multi_row_range(const pb_multi* m, int64_t n_rows, int rank, int64_t* lo, int64_t* hi);
int  pb_multi_sync(pb_multi* m);
int  pb_multi_pod(pb_multi* m, const double* const* U_dev, const int64_t* rows, int64_t n_st, const int64_t* ldu,
                  double eps, int n_L_in, int mode, int* n_L);
int  pb_multi_pod_host(pb_multi* m, const double* U_host, int64_t n_h, int64_t n_st, int64_t ldu_host,
                       double eps, int n_L_in, int mode, int* n_L);
/* two-step POD (pod.py:51-68) of row-sharded slabs with sample-major columns (n_st = n_t n_s); needs rows[p] >= n_t and
 * a tall packed basis (n_h >= sum r_k); pb_multi_basis afterwards */
int  pb_multi_fast_pod(pb_multi* m, const double* const* U_dev, const int64_t* rows, int n_t, int64_t n_s, const int64_t* ldu,
                       double eps, double eps_init, int mode, int* n_L, int* ranks_host);
int  pb_multi_basis(pb_multi* m, double* const* V_dev, const int64_t* ldv);
int  pb_multi_basis_host(pb_multi* m, double* V_host, int64_t ldv_host);
int  pb_multi_project(pb_multi* m, const double* const* V_dev, const int64_t* ldv, int n_L, const double* const* U_dev,
                      const int64_t* ldu, const int64_t* rows, int64_t n, double* const* v_dev);
int  pb_multi_project_host(pb_multi* m, double* v_host);
int  pb_multi_sigma(pb_multi* m, double* sigma_host, int64_t capacity, int64_t* count);
int  pb_multi_all_reduce(pb_multi* m, double* const* bufs, int64_t count);
int  pb_multi_timer_start(pb_multi* m);
int  pb_multi_timer_stop(pb_multi* m, double* ms);

/* ---- projection / reconstruction ----------------------------------------
 * replaces podnnmodel.py:435-437 (project_to_v), :431-433 (project_to_U), :247-253.
 *   pb_project:      v (n x n_L, ld = n_L) = (V^T U)^T for the local slab; partial sums
 *                    over the slab rows (all-reduce across ranks when sharded).
 *   pb_reconstruct:  U_pod (n_rows x n, ld) = V v^T; pod_sig (n_rows, nullable) =
 *                    mean_j |U - U_pod| / 2  (= np.stack((U,U_pod),-1).std(-1).mean(-1)).
 *                    U_pod_dev may be NULL when only pod_sig is wanted.
 */
int  pb_project(pb_ctx* ctx, const double* V_dev, int64_t ldv, int n_L,
                const double* U_dev, int64_t ldu, int64_t n_rows, int64_t n, double* v_dev);
int  pb_reconstruct(pb_ctx* ctx, const double* V_dev, int64_t ldv, int n_L,
                    const double* v_dev, int64_t n, int64_t n_rows,
                    const double* U_dev, int64_t ldu,
                    double* U_pod_dev, int64_t ldp, double* pod_sig_dev);

/* ---- deep-ensemble surrogate ---------------------------------------------
 * replaces varneuralnetwork.py:42-62,75-86,154-160 (member forward) and
 * podnnmodel.py:314-328 (predict_v mixture moments).
 *   dims[n_layers+1] = [n_d, h_1, ..., h_k, 2*n_L];   weights_dev = for each member,
 *   for each layer: W (in x out, row-major) then b (out), packed back to back.
 *   X_dev (N x n_d).  Outputs v_mean, v_sig (N x n_L).
 */
int  pb_ensemble_predict(pb_ctx* ctx, int n_members, int n_layers, const int* dims,
                         const double* weights_dev, int norm_kind,
                         const double* norm_a_dev, const double* norm_b_dev, double soft_0,
                         const double* X_dev, int64_t N, double* v_mean_dev, double* v_sig_dev);
/* U-level prediction, PodnnModel.predict (podnnmodel.py:366-379).
 * samples == 0: the exact moments of the reference's Monte-Carlo loop, U_mean = V v_mean^T,
 *               U_sig = sqrt((V*V) (v_sig^2)^T).
 * samples >= 2: the loop itself: `samples` draws v ~ N(v_mean, v_sig) from a Philox4x32-10 stream keyed
 *               by (seed, draw, point, mode), U_i = V v_i^T accumulated into running sum / sum of squares in
 *               the GEMM epilogue (U_i is never stored), then mean and the unbiased sample sigma (:376-378). */
int  pb_predict_U(pb_ctx* ctx, const double* V_dev, int64_t ldv, int64_t n_rows, int n_L,
                  const double* v_mean_dev, const double* v_sig_dev, int64_t N,
                  int samples, uint64_t seed, double* U_mean_dev, double* U_sig_dev, int64_t ldo);

/* Mixture of the ensemble members at the U level, PodnnModel.predict_mc (podnnmodel.py:344-362), kept in HBM:
 * accumulate: sum += U_mean_i, m2 += U_sig_i^2 + U_mean_i^2 (first != 0 overwrites instead of adding);
 * finalize (in place): sum <- U_pred = sum / n_members, m2 <- sqrt(m2 / n_members - U_pred^2) (+ pod_sig[row] when
 * pod_sig_dev is not NULL, :359-360).  All matrices n_rows x N with the given leading dimensions. */
int  pb_mixture_accumulate(pb_ctx* ctx, const double* U_mean_dev, const double* U_sig_dev, int64_t ld_in,
                           int64_t n_rows, int64_t N, double* sum_dev, double* m2_dev, int64_t ld_acc, int first);
int  pb_mixture_finalize(pb_ctx* ctx, double* sum_dev, double* m2_dev, int64_t ld, int64_t n_rows, int64_t N,
                         int n_members, const double* pod_sig_dev);

/* Ensemble training, SURVEY.md section 8 row f4: `epochs` full-batch steps of VarNeuralNetwork.grad +
 * tf_optimization_step (varneuralnetwork.py:94-128; optimiser tf.keras.optimizers.Adam(lr) :24, L2 term :88-92)
 * for all members at once.  dims = [n_d, hidden..., 2 n_L] (n_layers + 1 entries).
 * params_dev / adam_m_dev / adam_v_dev: n_members x P doubles, per member the Keras get_weights() order
 *   [W0 (in x out, row-major), b0, W1, b1, ...] flattened; updated in place (Adam moments start at zero,
 *   step0 = optimiser steps already taken, for the bias correction when a run is resumed).
 * use_adv != 0: the reference's `adv_eps is not None` branch (adv_eps = 0 included): X_adv = X + adv_eps
 *   sign(d loss / d X), second NLL + second L2 term.  X_dev (N x n_d) is the NORMALISED input (:136-137),
 *   v_dev (N x n_L) the targets, both shared by the members.
 * loss_host (epochs x n_members, may be NULL): loss before the update of each epoch, as tf_optimization_step
 *   returns it.  use_graph != 0 replays one captured CUDA graph per epoch (the default path); 0 launches eagerly. */
int  pb_ensemble_train(pb_ctx* ctx, int n_members, int n_layers, const int* dims, double* params_dev,
                       double* adam_m_dev, double* adam_v_dev, int64_t step0, int epochs, double lr, double lam,
                       int use_adv, double adv_eps, double soft_0, const double* X_dev, const double* v_dev,
                       int64_t N, double* loss_host, int use_graph);

/* ---- measurement helpers (used by bench.py and the tests; not on the product path) ------ */
/* Closed-form left basis of pb_dct_lowrank's matrix: Phi (n_rows x r, ld) rows [row0, row0 + n_rows) of
 * Phi[i][k] = sqrt(2/n_h) cos(pi (i + 1/2)(k + 1) / n_h)  (the exact modes of the C5 known-answer case). */
int  pb_dct_basis(pb_ctx* ctx, int64_t n_h, int r, int64_t row0, int64_t n_rows, double* Phi_dev, int64_t ld);
/* G_out (n_L x n_L) = R^T R with R = V - Phi C over the local rows (C: k x n_L, normally Phi^T V summed over all rows).
 * sqrt(lambda_max) of the sum over ranks = sin(largest principal angle between span V and span Phi). */
int  pb_residual_gram(pb_ctx* ctx, const double* V_dev, int64_t ldv, int n_L, const double* Phi_dev, int64_t ldp, int k,
                      int64_t n_rows, const double* C_dev, double* G_out_dev);
/* Host-only (no GPU needed): the warp work shapes of the Gram kernel's special tile kinds for a matrix whose last
 * 128-column tile holds `valid_atom_cols` (1..16) groups of 8 columns.  out[kind - 1][warp] = {r0, c0, nt} for
 * kind 1 = edge, 2 = diagonal, 3 = diagonal + edge: the warp accumulates the 8 x 8 atoms (r0 .. r0 + nt - 1) x
 * (c0 .. c0 + 3) of the 16 x 16-atom tile (of its transpose when transposed[kind - 1]); weight[kind - 1] = schedule
 * weight of one iteration: tenths of an atom of the busiest SM sub-partition + measured overhead (a full tile: 640). */
int  pb_gram_tile_shapes(int valid_atom_cols, int out[3][8][3], int transposed[3], int weight[3]);
/* Calibration of the Gram kernel's static schedule.  enable = 1: later pb_gram calls of this context record the busy time
 * of every persistent CTA.  With busy_us / iters (capacity >= the SM count) it returns, per CTA of the last Gram, that
 * time and the 16-row iterations it ran of each tile kind (full, edge, diagonal, diagonal + edge); n_cta = CTAs. */
int  pb_gram_profile(pb_ctx* ctx, int enable, int capacity, double* busy_us, int* iters, int* n_cta);
/* kind: 0 = DFMA register loop, 1 = DMMA m8n8k4 loop, 2 = both pipes mixed.  Returns TFLOP/s. */
int  pb_probe_fp64_peak(pb_ctx* ctx, int kind, double* tflops);
/* device-to-device copy of `bytes` (read+write counted).  Returns GB/s. */
int  pb_probe_hbm_copy(pb_ctx* ctx, size_t bytes, double* gbs);
/* GB/s of the host-side pageable -> pinned staging copy (threads <= 0 / streaming < 0: the library's own settings:
 * env PB_HOST_COPY_THREADS, PB_HOST_COPY_NT) */
int  pb_probe_host_copy(size_t bytes, int threads, int streaming, double* gbs);
/* write `bytes` of scratch to evict L2 between timed iterations */
int  pb_flush_l2(pb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PODB200_H */
