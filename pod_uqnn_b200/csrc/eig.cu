// Symmetric PSD eigensolver for the (small) Gram matrices of the POD:
//
//   G (m x m)  --pivoted Cholesky-->  L (m x r), G ~= L L^T      (rank revealing, optional
//                                                                  truncation / diagonal shift)
//   one-sided block Jacobi on the r column vectors of L (length m):  L J = W diag(sigma)
//   => eigenvalues lambda_k = sigma_k^2 (- shift), eigenvectors w_k = (L J)_k / sigma_k.
//
// Working on the Cholesky factor instead of G keeps the graded accuracy of the factor
// (errors relative to sqrt(lambda_i lambda_j), Demmel/Veselic; Drmac/Veselic preconditioning),
// costs O(m r^2) instead of O(m^3) when the snapshot set has low numerical rank r, and makes
// the Jacobi converge in a few sweeps.  The 32 x 32 pair problems are solved in shared memory
// by a two-sided cyclic Jacobi; the m-long Gram / update products of a pair run on DMMA.
//
// Vectors are stored one per row: At[k][0..ldm), ldm = m rounded up to 8, zero padded.
#include "common.cuh"
#include "tile_ops.cuh"
#include <cooperative_groups.h>
#include <cfloat>
#include <cstdlib>
#include <algorithm>
namespace cg = cooperative_groups;

namespace {

constexpr int JB = 16;            // vectors per block
constexpr int JP = 2 * JB;        // vectors per pair problem
constexpr int CHOL_ROWS = 32;     // rows per group in the Cholesky column update
constexpr int CHOL_THREADS = 256;

using pbt::dmma884;

// ---------------------------------------------------------------------------
// pivoted Cholesky, cooperative (one grid barrier per step)
// ---------------------------------------------------------------------------
struct CholArgs {
    const double* G; int m; int64_t ldg;
    double* Lt; int64_t ldm; int max_steps;
    double* d;              // (m) running Schur diagonal
    double* cand_val;       // (2 x grid) per-CTA arg-max candidates, double buffered
    int* cand_idx;
    double tol_rel; double shift_rel;
    int* r_out; double* shift_out;
    int k_start;            // > 0: continue a factorisation started by chol_small_kernel (d, shift, thresh in memory)
    const double* thresh_in;
};

__global__ void __launch_bounds__(CHOL_THREADS) chol_kernel(CholArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red_v[CHOL_THREADS];
    __shared__ int red_i[CHOL_THREADS];
    __shared__ double part[CHOL_THREADS / 32][CHOL_ROWS];
    extern __shared__ double lp[];            // (max_steps) row p of the factor so far
    const int tid = threadIdx.x, nth = blockDim.x;
    const int G_ = gridDim.x, cta = blockIdx.x;
    const int m = a.m;
    const int rows_per_cta = (m + G_ - 1) / G_;
    const int j_begin = cta * rows_per_cta, j_end = min(m, j_begin + rows_per_cta);

    // trace (every CTA redundantly, fixed order) -> shift
    double tr = 0.;
    for (int j = tid; j < m; j += nth) tr += a.G[(int64_t)j * a.ldg + j];
    red_v[tid] = tr; __syncthreads();
    for (int s = nth / 2; s > 0; s >>= 1) { if (tid < s) red_v[tid] += red_v[tid + s]; __syncthreads(); }
    const double shift = a.k_start > 0 ? *a.shift_out : a.shift_rel * red_v[0];
    __syncthreads();
    if (a.k_start == 0) for (int j = j_begin + tid; j < j_end; j += nth) a.d[j] = a.G[(int64_t)j * a.ldg + j] + shift;
    __syncthreads();

    double thresh = a.k_start > 0 ? *a.thresh_in : 0.;
    int k = a.k_start;
    for (; k < a.max_steps; ++k) {
        // local arg-max over my rows (ties -> lowest index)
        double bv = -DBL_MAX; int bi = 0x7fffffff;
        for (int j = j_begin + tid; j < j_end; j += nth) {
            double v = a.d[j];
            if (v > bv || (v == bv && j < bi)) { bv = v; bi = j; }
        }
        red_v[tid] = bv; red_i[tid] = bi; __syncthreads();
        for (int s = nth / 2; s > 0; s >>= 1) {
            if (tid < s) {
                double v = red_v[tid + s]; int i = red_i[tid + s];
                if (v > red_v[tid] || (v == red_v[tid] && i < red_i[tid])) { red_v[tid] = v; red_i[tid] = i; }
            }
            __syncthreads();
        }
        const int buf = (k & 1) * G_;
        if (tid == 0) { a.cand_val[buf + cta] = red_v[0]; a.cand_idx[buf + cta] = red_i[0]; }
        grid.sync();
        // every CTA reduces the candidates identically
        bv = -DBL_MAX; bi = 0x7fffffff;
        for (int c = tid; c < G_; c += nth) {
            double v = a.cand_val[buf + c]; int i = a.cand_idx[buf + c];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
        __syncthreads();
        red_v[tid] = bv; red_i[tid] = bi; __syncthreads();
        for (int s = nth / 2; s > 0; s >>= 1) {
            if (tid < s) {
                double v = red_v[tid + s]; int i = red_i[tid + s];
                if (v > red_v[tid] || (v == red_v[tid] && i < red_i[tid])) { red_v[tid] = v; red_i[tid] = i; }
            }
            __syncthreads();
        }
        const double dp = red_v[0]; const int p = red_i[0];
        __syncthreads();
        if (k == 0) thresh = a.tol_rel * dp;
        if (!(dp > thresh) || !(dp > 0.)) break;          // uniform across the grid
        (void)0;
        // row p of the factor so far (written by the owner of p in earlier steps)
        for (int q = tid; q < k; q += nth) lp[q] = a.Lt[(int64_t)q * a.ldm + p];
        __syncthreads();
        const double inv = 1. / sqrt(dp);
        // column k: v_j = (G[j][p] (+shift) - sum_q L[j][q] L[p][q]) / sqrt(dp)
        const int jr = tid % CHOL_ROWS, qpart = tid / CHOL_ROWS;
        constexpr int QP = CHOL_THREADS / CHOL_ROWS;
        for (int jg = j_begin; jg < j_end; jg += CHOL_ROWS) {
            const int j = jg + jr;
            double s = 0.;
            if (j < j_end) {
                const double* col = a.Lt + j;
                int q = qpart;
                double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
                for (; q + 3 * QP < k; q += 4 * QP) {
                    s0 += col[(int64_t)q * a.ldm] * lp[q];
                    s1 += col[(int64_t)(q + QP) * a.ldm] * lp[q + QP];
                    s2 += col[(int64_t)(q + 2 * QP) * a.ldm] * lp[q + 2 * QP];
                    s3 += col[(int64_t)(q + 3 * QP) * a.ldm] * lp[q + 3 * QP];
                }
                for (; q < k; q += QP) s0 += col[(int64_t)q * a.ldm] * lp[q];
                s = (s0 + s1) + (s2 + s3);
            }
            part[qpart][jr] = s;
            __syncthreads();
            if (qpart == 0 && j < j_end) {
                double dot = 0.;
#pragma unroll
                for (int w = 0; w < QP; ++w) dot += part[w][jr];
                double g = a.G[(int64_t)j * a.ldg + p] + (j == p ? shift : 0.);
                double v = (g - dot) * inv;
                if (a.d[j] == -DBL_MAX) v = 0.;             // rows of earlier pivots stay zero
                if (j == p) v = sqrt(dp);
                a.Lt[(int64_t)k * a.ldm + j] = v;
                a.d[j] = (j == p || a.d[j] == -DBL_MAX) ? -DBL_MAX : a.d[j] - v * v;
            }
            __syncthreads();
        }
    }
    if (cta == 0 && tid == 0) { *a.r_out = k; *a.shift_out = shift; }
}

// ---------------------------------------------------------------------------
// one-sided block Jacobi: one CTA per block pair per round
// ---------------------------------------------------------------------------
__device__ __forceinline__ void rr_pair(int n, int round, int i, int& p, int& q) {
    // circle method on n (even) players: player n-1 is fixed
    if (i == 0) { p = n - 1; q = round % (n - 1); }
    else { p = (round + i) % (n - 1); q = (round - i + 2 * (n - 1)) % (n - 1); }
    if (p > q) { int t = p; p = q; q = t; }
}

constexpr int SP = JP + 1;   // padded pitch of the 32 x 32 shared matrices

// Two-sided cyclic Jacobi on S (JP x JP, symmetric, in shared memory), accumulating the
// rotations in Q.  256 threads: thread (P, R) owns the 2 x 2 block rows(pair P) x cols(pair R).
// Returns the number of rotations applied in the FIRST sweep (0 => the pair was converged).
// pair `idx` (0..15) of inner round `round`: full mode = round-robin tournament over all 32 vectors
// (31 rounds); cross mode = vector idx of the first block against vector (idx + round) % 16 of the second
// (16 rounds cover the 256 cross pairs exactly once)
__device__ __forceinline__ void inner_pair(int cross, int round, int idx, int& p, int& q) {
    if (cross) { p = idx; q = JB + ((idx + round) & (JB - 1)); }
    else rr_pair(JP, round, idx, p, q);
}

// Two-sided Jacobi on S (JP x JP, symmetric, in shared memory), accumulating the rotations in Q.
// 256 threads: thread (P, R) owns the 2 x 2 block rows(pair P) x cols(pair R).
//   cross = 0: cyclic sweeps over all pairs until a sweep does nothing (at most max_inner sweeps);
//   cross = 1: ONE pass over the 256 pairs between the two 16-vector blocks -- the classical block-cyclic
//              one-sided Jacobi, where every column pair is rotated once per outer sweep.  The rounds are
//              barrier-separated and latency bound (16 threads compute the rotations, 240 wait), so 16
//              rounds instead of up to 3 x 31 is what makes a round of the outer sweep cheap.
// Returns the number of rotations applied in the first pass (0 => nothing to do for this pair).
__device__ int inner_jacobi(double (*S)[SP], double (*Q)[SP], double* cs, double* sn, double tol, int max_inner,
                            int cross, int* clean_exit) {
    const int tid = threadIdx.x;
    for (int e = tid; e < JP * JP; e += blockDim.x) Q[e / JP][e % JP] = (e / JP == e % JP) ? 1. : 0.;
    __syncthreads();
    int first = 0;
    const int n_rounds = cross ? JB : JP - 1;
    const int n_sweeps = cross ? 1 : max_inner;
    for (int sweep = 0; sweep < n_sweeps; ++sweep) {
        int rot_sweep = 0;
        for (int round = 0; round < n_rounds; ++round) {
            int did = 0;
            if (tid < JP / 2) {
                int p, q; inner_pair(cross, round, tid, p, q);
                double app = S[p][p], aqq = S[q][q], apq = S[p][q];
                double c = 1., s = 0.;
                if (apq * apq > tol * tol * fabs(app * aqq) && apq != 0.) {
                    // small-angle Jacobi rotation from cos(2 theta) = |d| / h, h = hypot(d, 2 apq); two
                    // reciprocal square roots instead of two divisions and two square roots (the round's
                    // critical path: 240 threads wait for these 16)
                    const double d = aqq - app;
                    const double rh = rsqrt(d * d + 4. * apq * apq);
                    const double c2 = 0.5 + 0.5 * fabs(d) * rh;          // cos^2 theta in [1/2, 1]
                    const double rc = rsqrt(c2);
                    c = c2 * rc;
                    s = (d >= 0. ? apq : -apq) * rh * rc;                 // sin theta, sign of zeta = d / (2 apq)
                    did = 1;
                }
                cs[tid] = c; sn[tid] = s;
            }
            const int rot_round = __syncthreads_count(did);
            rot_sweep += rot_round;
            if (rot_round == 0) continue;      // no pair of this round rotates (late sweeps): skip the apply pass and its barrier
            {
                const int P = tid / (JP / 2), R = tid % (JP / 2);
                int p1, q1, p2, q2; inner_pair(cross, round, P, p1, q1); inner_pair(cross, round, R, p2, q2);
                const double c1 = cs[P], s1 = sn[P], c2 = cs[R], s2 = sn[R];
                double m11 = S[p1][p2], m12 = S[p1][q2], m21 = S[q1][p2], m22 = S[q1][q2];
                // right: columns (p2, q2) <- (c2 col_p - s2 col_q, s2 col_p + c2 col_q)
                double n11 = c2 * m11 - s2 * m12, n12 = s2 * m11 + c2 * m12;
                double n21 = c2 * m21 - s2 * m22, n22 = s2 * m21 + c2 * m22;
                // left: rows (p1, q1) <- (c1 row_p - s1 row_q, s1 row_p + c1 row_q)
                S[p1][p2] = c1 * n11 - s1 * n21; S[p1][q2] = c1 * n12 - s1 * n22;
                S[q1][p2] = s1 * n11 + c1 * n21; S[q1][q2] = s1 * n12 + c1 * n22;
                // Q <- Q J : rows split over the two half-blocks of threads
                for (int row = P; row < JP; row += JP / 2) {
                    double a = Q[row][p2], b = Q[row][q2];
                    Q[row][p2] = c2 * a - s2 * b; Q[row][q2] = s2 * a + c2 * b;
                }
            }
            __syncthreads();
        }
        if (sweep == 0) first = rot_sweep;
        *clean_exit = (rot_sweep == 0);       // the sweeps ended on one without rotations: converged to the threshold
        if (rot_sweep == 0) break;
    }
    return first;
}

struct JacArgs {
    double* At; int64_t ldm; int nb; int round; double tol; int* n_rot; int max_inner; int64_t batch_stride; int cross;
    int* clean;      // single-pair problems: set to 1 when the in-kernel solve ended on a sweep without rotations (nullable)
};

constexpr int JCH = 256;        // vector elements staged per shared-memory chunk
constexpr int JPX = JCH + 4;    // chunk pitch (== 4 mod 16: conflict-free DMMA fragment reads)

// stage elements [c0, c0 + JCH) of the 32 pair vectors into Xs[32][JPX] (zero beyond ldm)
__device__ __forceinline__ void stage_chunk(double* Xs, const double* At, int64_t ldm, int bp, int bq,
                                            int64_t c0, int tid) {
#pragma unroll
    for (int q = 0; q < (JP * JCH / 2) / 256; ++q) {
        const int idx = tid + q * 256;
        const int v = idx / (JCH / 2), c = (idx % (JCH / 2)) * 2;
        const double* row = At + (int64_t)((v < JB ? bp : bq) * JB + (v % JB)) * ldm;
        const int64_t e = c0 + c;
        const int bytes = e < ldm ? 16 : 0;
        pbt::cp_async16(Xs + v * JPX + c, bytes ? row + e : At, bytes);
    }
}

__global__ void __launch_bounds__(256) jacobi_pair_kernel(JacArgs a) {
    __shared__ double S[JP][SP];
    __shared__ double Q[JP][SP];
    __shared__ double cs[JP / 2], sn[JP / 2];
    extern __shared__ __align__(16) double dyn[];      // two chunk buffers; later the 8 partial Grams
    double* Xs[2] = {dyn, dyn + JP * JPX};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lk = lane & 3, lg = lane >> 2;
    int bp, bq; rr_pair(a.nb, a.round, blockIdx.x, bp, bq);
    a.At += (int64_t)blockIdx.y * a.batch_stride;            // batch of independent factors
    // vector v (0..31) of the pair lives at row: (v < 16 ? bp : bq) * 16 + v % 16
    auto vrow = [&](int v) -> double* {
        return a.At + (int64_t)((v < JB ? bp : bq) * JB + (v % JB)) * a.ldm;
    };
    const int nch = (int)((a.ldm + JCH - 1) / JCH);

    // ---- phase 1: S = X X^T over the m-long vectors (DMMA; the fragment of row block t is A and B)
    {
        double acc[4][4][2];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }
        stage_chunk(Xs[0], a.At, a.ldm, bp, bq, 0, tid);
        pbt::cp_async_commit();
        for (int c = 0; c < nch; ++c) {
            if (c + 1 < nch) stage_chunk(Xs[(c + 1) & 1], a.At, a.ldm, bp, bq, (int64_t)(c + 1) * JCH, tid);
            pbt::cp_async_commit();
            pbt::cp_async_wait<1>();
            __syncthreads();
            const double* x = Xs[c & 1] + lg * JPX + lk;
#pragma unroll 2
            for (int k4 = warp; k4 < JCH / 4; k4 += 8) {
                double f[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) f[t] = x[t * 8 * JPX + k4 * 4];
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) dmma884(acc[t][u][0], acc[t][u][1], f[t], f[u]);
            }
            __syncthreads();
        }
        pbt::cp_async_wait<0>();
        double* red = dyn;                                  // 8 x (JP*JP) doubles, fits the first buffer
        double* my = red + warp * (JP * JP);
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                my[(t * 8 + lg) * JP + u * 8 + 2 * lk] = acc[t][u][0];
                my[(t * 8 + lg) * JP + u * 8 + 2 * lk + 1] = acc[t][u][1];
            }
        __syncthreads();
        for (int e = tid; e < JP * JP; e += blockDim.x) {
            double s = 0.;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w * (JP * JP) + e];
            S[e / JP][e % JP] = s;
        }
        __syncthreads();
        // exact symmetry (the two triangles were accumulated in different lanes)
        for (int e = tid; e < JP * JP; e += blockDim.x) {
            int i = e / JP, j = e % JP;
            if (j < i) S[i][j] = S[j][i];
        }
        __syncthreads();
    }

    // ---- phase 2: small eigenproblem
    // round 0 of a sweep meets every block exactly once: it also rotates the pairs INSIDE the two blocks;
    // the other rounds only the 256 cross pairs (a.cross)
    int clean = 0;
    const int first = inner_jacobi(S, Q, cs, sn, a.tol, a.max_inner, a.cross, &clean);
    if (a.clean && tid == 0) *a.clean = clean;
    if (first == 0) return;                       // uniform: nothing to rotate
    if (tid == 0) atomicAdd(a.n_rot, first);

    // ---- phase 3: X' = Q^T X, chunk by chunk through shared memory, written back in place
    {
        // A fragments of Q^T: A[row = j][k = i] = Q[i][j]
        double qa[4][8];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) qa[t][ks] = Q[ks * 4 + lk][t * 8 + lg];
        double* dst[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) dst[t] = vrow(t * 8 + lg) + 2 * lk;
        __syncthreads();
        stage_chunk(Xs[0], a.At, a.ldm, bp, bq, 0, tid);
        pbt::cp_async_commit();
        for (int c = 0; c < nch; ++c) {
            if (c + 1 < nch) stage_chunk(Xs[(c + 1) & 1], a.At, a.ldm, bp, bq, (int64_t)(c + 1) * JCH, tid);
            pbt::cp_async_commit();
            pbt::cp_async_wait<1>();
            __syncthreads();
            const double* x = Xs[c & 1] + lk * JPX + lg;
            for (int e8 = warp; e8 < JCH / 8; e8 += 8) {
                const int64_t e = (int64_t)c * JCH + e8 * 8;
                if (e >= a.ldm) break;
                double b[8];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) b[ks] = x[ks * 4 * JPX + e8 * 8];
                double d[4][2];
#pragma unroll
                for (int t = 0; t < 4; ++t) { d[t][0] = 0.; d[t][1] = 0.; }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                    for (int t = 0; t < 4; ++t) dmma884(d[t][0], d[t][1], qa[t][ks], b[ks]);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    *reinterpret_cast<double2*>(dst[t] + e) = make_double2(d[t][0], d[t][1]);
            }
            __syncthreads();
        }
        pbt::cp_async_wait<0>();
    }
}

// squared norms -> sigma, then sort descending (single CTA bitonic sort)
__global__ void norms_kernel(const double* __restrict__ At, int64_t ldm, int nvec, double* __restrict__ nrm2,
                             int64_t bsAt, int64_t bsN) {
    const int v = blockIdx.x;
    if (v >= nvec) return;
    At += (int64_t)blockIdx.y * bsAt; nrm2 += (int64_t)blockIdx.y * bsN;
    __shared__ double red[256];
    double s = 0.;
    for (int64_t e = threadIdx.x; e < ldm; e += blockDim.x) { double x = At[(int64_t)v * ldm + e]; s += x * x; }
    red[threadIdx.x] = s; __syncthreads();
    for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
    if (threadIdx.x == 0) nrm2[v] = red[0];
}

__global__ void sort_kernel(const double* __restrict__ nrm2, int nvec, int npow2, const double* shift,
                            double* __restrict__ sig_out, double* __restrict__ nrm_out, int m,
                            int* __restrict__ perm_out, int* r_eff, int64_t bsN, int64_t bsS, int64_t bsP) {
    nrm2 += (int64_t)blockIdx.x * bsN; shift += blockIdx.x * (bsN ? 1 : 0);
    sig_out += (int64_t)blockIdx.x * bsS; nrm_out += (int64_t)blockIdx.x * bsS;
    perm_out += (int64_t)blockIdx.x * bsP; r_eff += blockIdx.x * (bsN ? 1 : 0);
    extern __shared__ unsigned char sm[];
    double* key = reinterpret_cast<double*>(sm);
    int* idx = reinterpret_cast<int*>(sm + sizeof(double) * npow2);
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        key[i] = i < nvec ? nrm2[i] : -1.; idx[i] = i;
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
                int l = i ^ j;
                if (l > i) {
                    bool desc = ((i & k) == 0);
                    // order: larger key first; ties -> smaller index first (deterministic)
                    bool i_first = key[i] > key[l] || (key[i] == key[l] && idx[i] < idx[l]);
                    if (desc ? !i_first : i_first) {
                        double tk = key[i]; key[i] = key[l]; key[l] = tk;
                        int ti = idx[i]; idx[i] = idx[l]; idx[l] = ti;
                    }
                }
            }
            __syncthreads();
        }
    const double sh = *shift;
    int cnt = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        double lam = (i < nvec && key[i] > 0.) ? key[i] - sh : 0.;
        double s = lam > 0. ? sqrt(lam) : 0.;
        sig_out[i] = s;
        nrm_out[i] = (i < nvec && key[i] > 0.) ? sqrt(key[i]) : 0.;
        if (i < nvec) perm_out[i] = idx[i];
        if (s > 0.) ++cnt;
    }
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    atomicAdd(&total, cnt);
    __syncthreads();
    if (threadIdx.x == 0) *r_eff = total;
}

// truncation rank (pod.py:22-32): smallest n_L with cumsum(lambda)[n_L-1] / sum(lambda) >= 1 - eps.
// One warp; each lane accumulates a contiguous chunk in order, a shuffle scan chains the chunks.
__global__ void rank_kernel(const double* __restrict__ sig, int count, double eps, int* n_L_out, int64_t bsS) {
    sig += (int64_t)blockIdx.x * bsS; n_L_out += blockIdx.x;
    const int lane = threadIdx.x;
    const int chunk = (count + 31) / 32;
    const int b = lane * chunk, e = min(count, b + chunk);
    double s = 0.;
    for (int i = b; i < e; ++i) s += sig[i] * sig[i];
    double incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const double total = __shfl_sync(0xffffffffu, incl, 31);
    double acc = incl - s;
    const double thr = 1. - eps;
    int found = count;             // loop ran out: every mode kept
    for (int i = b; i < e; ++i) {
        acc += sig[i] * sig[i];
        if (acc / total >= thr) { found = i + 1; break; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
    if (lane == 0) *n_L_out = found;
}

__global__ void factor_to_right_kernel(const double* __restrict__ At, int64_t ldm, const int* __restrict__ perm,
                                       const double* __restrict__ nrm, const double* __restrict__ sig,
                                       int m, int k, double* __restrict__ B, int64_t ldb,
                                       int64_t bsAt, int64_t bsS, int64_t bsP, int64_t bsB) {
    At += (int64_t)blockIdx.y * bsAt; perm += (int64_t)blockIdx.y * bsP; nrm += (int64_t)blockIdx.y * bsS;
    if (sig) sig += (int64_t)blockIdx.y * bsS;
    B += (int64_t)blockIdx.y * bsB;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)m * k) return;
    int i = (int)(idx / k), j = (int)(idx % k);
    double d = nrm[j];
    if (sig) d *= sig[j];
    B[(int64_t)i * ldb + j] = d > 0. ? At[(int64_t)perm[j] * ldm + i] / d : 0.;
}

}  // namespace

static int g_cross_rounds = []{ const char* e = getenv("PB_JACOBI_CROSS"); return e ? atoi(e) : 1; }();
static int g_inner_sweeps = []{ const char* e = getenv("PB_JACOBI_INNER"); return e ? atoi(e) : 3; }();
// device ints layout in ctx->d_int: [0] r (Cholesky), [1] n_rot, [2] n_L, [3] r_eff
struct EigScratch { double* d; double* cand_val; int* cand_idx; double* nrm2; double* shift; };

static int get_scratch(pb_ctx* ctx, int64_t m, EigScratch* s) {
    // small: [d (m)] [cand_val (2*1024)] [nrm2 (m+32)] [shift (2)] then ints [cand_idx (2*1024)]
    size_t dbl = (size_t)m + 2048 + (m + 32) + 2;
    size_t ints = 2048;
    PB_TRY(pb_reserve(ctx, &ctx->small, &ctx->small_cap, dbl + (ints + 1) / 2 + 8));
    s->d = ctx->small; s->cand_val = s->d + m; s->nrm2 = s->cand_val + 2048; s->shift = s->nrm2 + (m + 32);
    s->cand_idx = reinterpret_cast<int*>(s->shift + 2);
    return PB_OK;
}

// One-sided block Jacobi on the r factor vectors held in At (rows of length ldm, zero padded up to a multiple of 32
// rows), then norms, sort, sigma = sqrt(norm^2 - *shift_dev).  Shared by the Cholesky path (pbi_eig_psd) and the TSQR
// path (qr.cu: rows of the pivoted R factor).
int pbi_jacobi_factor(pb_ctx* ctx, double* At, int64_t m, int r, const double* shift_dev, double* sig_dev, double* nrm_dev,
                      int* perm_dev, int* sweeps_out, int max_sweeps) {
    const int64_t ldm = (m + 7) / 8 * 8;
    EigScratch sc; PB_TRY(get_scratch(ctx, m, &sc));
    double* nrm2_scr = sc.nrm2;
    if (!shift_dev) {                                         // no diagonal shift (TSQR path)
        PB_CUDA(ctx, cudaMemsetAsync(sc.shift, 0, sizeof(double), ctx->stream));
        shift_dev = sc.shift;
    }
    // ---- one-sided block Jacobi on the r vectors
    int nb = (r + JB - 1) / JB; if (nb & 1) ++nb;             // even number of blocks (zero padded)
    const int nvec = nb * JB;
    // rotation threshold relative to the column norms (dgesvj uses sqrt(m) eps): below it the
    // pair Gram entry is rounding noise of the m-long dot product
    const double tol = ctx->eig_jacobi_tol > 0. ? ctx->eig_jacobi_tol : std::max(4., sqrt((double)m)) * DBL_EPSILON;
    ctx->eig_jacobi_tol = 0.;
    size_t jac_smem = sizeof(double) * 2 * JP * JPX;
    PB_TRY(pbi_smem_optin(ctx, (const void*)jacobi_pair_kernel, (size_t)(jac_smem)));
    int sweeps = 0; bool converged = false;
    const bool trust_single = ctx->eig_trust_single_pair && nb == 2;
    ctx->eig_trust_single_pair = false;
    for (; sweeps < max_sweeps && !converged; ++sweeps) {
        PB_CUDA(ctx, cudaMemsetAsync(ctx->d_int + 1, 0, sizeof(int), ctx->stream));
        const int rounds = nb > 2 ? nb - 1 : 1;
        for (int round = 0; round < rounds; ++round) {
            // a single pair (r <= 32): its 32 x 32 problem IS the whole problem -- solve it to convergence at once
            JacArgs ja{At, ldm, nb, round, tol, ctx->d_int + 1, nb == 2 ? 16 : g_inner_sweeps, 0, (g_cross_rounds && round > 0) ? 1 : 0,
                       nb == 2 ? ctx->d_int + 4 : nullptr};
            jacobi_pair_kernel<<<nb / 2, 256, jac_smem, ctx->stream>>>(ja);
            PB_LAUNCH_CHECK(ctx);
        }
        if (trust_single) { converged = true; continue; }
        // d_int[1] = rotations of the first inner sweep, d_int[4] = the single-pair solve ended on a clean sweep
        PB_CUDA(ctx, cudaMemcpyAsync(ctx->h_int + 1, ctx->d_int + 1, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        // a single pair iterated to a clean sweep inside the kernel IS converged: no confirming launch (82 us + a sync at C2)
        converged = (ctx->h_int[1] == 0) || (nb == 2 && ctx->h_int[4] == 1);
    }
    if (!converged && max_sweeps >= 30)
        PB_FAIL(ctx, PB_ERR_NOT_CONVERGED, "eig_psd: Jacobi did not converge in %d sweeps (m=%lld r=%d)", max_sweeps, (long long)m, r);

    // ---- norms, sort, sigma
    norms_kernel<<<nvec, 256, 0, ctx->stream>>>(At, ldm, nvec, nrm2_scr, 0, 0);
    PB_LAUNCH_CHECK(ctx);
    int npow2 = 1; while (npow2 < nvec) npow2 <<= 1;
    size_t sort_smem = (sizeof(double) + sizeof(int)) * (size_t)npow2;
    PB_TRY(pbi_smem_optin(ctx, (const void*)sort_kernel, (size_t)(sort_smem)));
    sort_kernel<<<1, 1024, sort_smem, ctx->stream>>>(nrm2_scr, nvec, npow2, shift_dev, sig_dev, nrm_dev, (int)m, perm_dev, ctx->d_int + 3, 0, 0, 0);
    PB_LAUNCH_CHECK(ctx);
    if (sweeps_out) *sweeps_out = sweeps;
    return PB_OK;
}


int pbi_eig_psd(pb_ctx* ctx, const double* G, int64_t m, double tol_rel, double shift_rel,
                double** At_buf, size_t* At_cap, double* sig_dev, double* nrm_dev, int* perm_dev,
                int* r_out, int* sweeps_out, int max_sweeps) {
    if (m <= 0 || m > 8192) PB_FAIL(ctx, PB_ERR_SHAPE, "eig_psd: order %lld outside (0, 8192]", (long long)m);
    const int64_t ldm = (m + 7) / 8 * 8;
    const int64_t cap_rows = (m + JP - 1) / JP * JP;          // even number of 16-blocks
    PB_TRY(pb_reserve(ctx, At_buf, At_cap, (size_t)cap_rows * ldm));
    double* At = *At_buf;
    EigScratch sc; PB_TRY(get_scratch(ctx, m, &sc));
    PB_CUDA(ctx, cudaMemsetAsync(At, 0, sizeof(double) * (size_t)cap_rows * ldm, ctx->stream));

    // ---- pivoted Cholesky.  Low numerical rank is the common case (the snapshot sets POD is used on): the first
    // SMALL_STEPS steps run in ONE CTA without grid barriers (~1.5 us per step instead of ~7); if the pivots have
    // not dropped below the tolerance by then, the cooperative multi-CTA kernel resumes from that state.
    constexpr int SMALL_STEPS = 64;
    int k_start = 0;
    if (m <= 1024) {
        // chol_small_kernel is declared further down in this file
        extern int pbi_chol_small_launch(pb_ctx*, const double*, int64_t, double*, int64_t, double, double, int*, double*, int,
                                         double*, double*, int*);
        PB_TRY(pbi_chol_small_launch(ctx, G, m, At, ldm, tol_rel, shift_rel, ctx->d_int + 0, sc.shift, SMALL_STEPS, sc.d,
                                     sc.shift + 1, ctx->d_int + 5));
        PB_CUDA(ctx, cudaMemcpyAsync(ctx->h_int, ctx->d_int, 6 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        k_start = ctx->h_int[5] ? -1 : ctx->h_int[0];      // -1: finished
    }
    if (k_start >= 0) {
        int grid = (int)std::min<int64_t>(ctx->sm_count, (m + CHOL_ROWS - 1) / CHOL_ROWS);
        size_t chol_smem = sizeof(double) * (size_t)m;
        PB_TRY(pbi_smem_optin(ctx, (const void*)chol_kernel, (size_t)(chol_smem)));
        int occ = 0;
        PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, chol_kernel, CHOL_THREADS, chol_smem));
        if (occ < 1) PB_FAIL(ctx, PB_ERR_CUDA, "eig_psd: Cholesky kernel does not fit an SM");
        grid = std::min(grid, occ * ctx->sm_count);
        if (grid > 1024) grid = 1024;
        CholArgs ca{G, (int)m, m, At, ldm, (int)m, sc.d, sc.cand_val, sc.cand_idx, tol_rel, shift_rel,
                    ctx->d_int + 0, sc.shift, k_start, sc.shift + 1};
        void* args[] = {&ca};
        PB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)chol_kernel, dim3(grid), dim3(CHOL_THREADS), args, chol_smem, ctx->stream));
        ctx->launches++;
        PB_CUDA(ctx, cudaMemcpyAsync(ctx->h_int, ctx->d_int, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const int r = ctx->h_int[0];
    if (r <= 0) PB_FAIL(ctx, PB_ERR_SHAPE, "eig_psd: matrix has no positive pivot (zero or indefinite input)");
    PB_TRY(pbi_jacobi_factor(ctx, At, m, r, sc.shift, sig_dev, nrm_dev, perm_dev, sweeps_out, max_sweeps));
    *r_out = r;                                   // factor vectors; sig may be zero for shifted noise modes
    return PB_OK;
}

// Orders <= 32 (the second-pass Gram of the rotated block, C2: 27): everything a decomposition needs -- pivoted Cholesky
// in one CTA, the single-pair Jacobi solve iterated to a clean sweep inside its kernel, norms, sort, the rank rule --
// is queued WITHOUT a host round trip; the caller synchronises once and reads h_int: [0] r, [2] n_L, [4] the Jacobi
// solve ended on a clean sweep, [5] the Cholesky finished.  (pbi_eig_psd + pbi_rank take three round trips for the same
// kernels: the factor rank, the convergence flag, the rank.)  If a flag is not as expected the caller repeats the
// decomposition through pbi_eig_psd; G is not modified.
int pbi_chol_small_launch(pb_ctx*, const double*, int64_t, double*, int64_t, double, double, int*, double*, int, double*, double*, int*);
int pbi_eig_psd_small_queued(pb_ctx* ctx, const double* G, int64_t m, double** At_buf, size_t* At_cap, double* sig_dev,
                             double* nrm_dev, int* perm_dev, double eps) {
    if (m <= 0 || m > JP) return 1;
    const int64_t ldm = (m + 7) / 8 * 8;
    const int64_t cap_rows = JP;
    PB_TRY(pb_reserve(ctx, At_buf, At_cap, (size_t)cap_rows * ldm));
    double* At = *At_buf;
    EigScratch sc; PB_TRY(get_scratch(ctx, m, &sc));
    PB_CUDA(ctx, cudaMemsetAsync(At, 0, sizeof(double) * (size_t)cap_rows * ldm, ctx->stream));
    PB_CUDA(ctx, cudaMemsetAsync(ctx->d_int, 0, 6 * sizeof(int), ctx->stream));
    PB_TRY(pbi_chol_small_launch(ctx, G, m, At, ldm, 0., 0., ctx->d_int + 0, sc.shift, (int)m, sc.d, sc.shift + 1, ctx->d_int + 5));
    const double tol = std::max(4., sqrt((double)m)) * DBL_EPSILON;
    const size_t jac_smem = sizeof(double) * 2 * JP * JPX;
    PB_TRY(pbi_smem_optin(ctx, (const void*)jacobi_pair_kernel, jac_smem));
    JacArgs ja{At, ldm, 2, 0, tol, ctx->d_int + 1, 16, 0, 0, ctx->d_int + 4};
    jacobi_pair_kernel<<<1, 256, jac_smem, ctx->stream>>>(ja);
    PB_LAUNCH_CHECK(ctx);
    const int nvec = JP;
    norms_kernel<<<nvec, 256, 0, ctx->stream>>>(At, ldm, nvec, sc.nrm2, 0, 0);
    PB_LAUNCH_CHECK(ctx);
    const int npow2 = JP;
    const size_t sort_smem = (sizeof(double) + sizeof(int)) * (size_t)npow2;
    PB_TRY(pbi_smem_optin(ctx, (const void*)sort_kernel, sort_smem));
    sort_kernel<<<1, 1024, sort_smem, ctx->stream>>>(sc.nrm2, nvec, npow2, sc.shift, sig_dev, nrm_dev, (int)m, perm_dev, ctx->d_int + 3, 0, 0, 0);
    PB_LAUNCH_CHECK(ctx);
    rank_kernel<<<1, 32, 0, ctx->stream>>>(sig_dev, (int)m, eps, ctx->d_int + 2, 0);
    PB_LAUNCH_CHECK(ctx);
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->h_int, ctx->d_int, 6 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

int pbi_rank(pb_ctx* ctx, const double* sig_dev, int64_t count, double eps, int* n_L_out) {
    rank_kernel<<<1, 32, 0, ctx->stream>>>(sig_dev, (int)count, eps, ctx->d_int + 2, 0);
    PB_LAUNCH_CHECK(ctx);
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->h_int + 2, ctx->d_int + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_L_out = ctx->h_int[2];
    return PB_OK;
}

int pbi_factor_to_right(pb_ctx* ctx, const double* At, int64_t m, int k, const int* perm_dev,
                        const double* nrm_dev, const double* sig_dev, double* B, int64_t ldb) {
    const int64_t ldm = (m + 7) / 8 * 8;
    int64_t n = m * k;
    factor_to_right_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(At, ldm, perm_dev, nrm_dev, sig_dev, (int)m, k, B, ldb, 0, 0, 0, 0);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}


// ---------------------------------------------------------------------------
// batched variant (two-step POD, pod.py:51-68): `batch` independent small PSD matrices
// ---------------------------------------------------------------------------
namespace {
// warp arg-max over (v, i) for v > 0 (anything else counts as 0): on return every lane holds the largest v and the
// smallest index that carries it; (0., 0x7fffffff) when no lane has a positive value
__device__ __forceinline__ void warp_argmax_pos(double& v, int& i) {
    const unsigned long long key = v > 0. ? (unsigned long long)__double_as_longlong(v) : 0ull;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const bool win = (hi == mh) && (lo == ml) && key != 0ull;
    const unsigned mi = __reduce_min_sync(0xffffffffu, win ? (unsigned)i : 0x7fffffffu);
    v = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
    i = (int)mi;
}

// pivoted Cholesky of one matrix per CTA (m <= 1024); same arithmetic as chol_kernel
__global__ void __launch_bounds__(1024) chol_small_kernel(const double* __restrict__ Gall, int m, double* __restrict__ Atall,
                                                         int64_t ldm, int64_t bsAt, double tol_rel, double shift_rel,
                                                         int* __restrict__ r_out, double* __restrict__ shift_out,
                                                         int max_steps, double* __restrict__ d_out,
                                                         double* __restrict__ thresh_out, int* __restrict__ done_out, int srows) {
    extern __shared__ double chol_sm[];          // d[m], lp[m], Ls[srows][m]: the first srows factor rows stay in shared memory
    double* d = chol_sm; double* lp = chol_sm + m; double* Ls = chol_sm + 2 * m;
    __shared__ double red_v[1024];
    __shared__ int red_i[1024];
    const int nth = blockDim.x, nwarps = blockDim.x >> 5;
    const double* G = Gall + (int64_t)blockIdx.x * m * m;
    double* At = Atall + (int64_t)blockIdx.x * bsAt;
    const int tid = threadIdx.x;
    double tr = 0.;
    for (int j = tid; j < m; j += nth) tr += G[(int64_t)j * m + j];
    red_v[tid] = tr; __syncthreads();
    for (int s = nth / 2; s > 0; s >>= 1) { if (tid < s) red_v[tid] += red_v[tid + s]; __syncthreads(); }
    const double shift = shift_rel * red_v[0];
    __syncthreads();
    for (int j = tid; j < m; j += nth) d[j] = G[(int64_t)j * m + j] + shift;
    __syncthreads();
    double thresh = 0.;
    int k = 0, done = 0;
    for (; k < max_steps; ++k) {
        double bv = -DBL_MAX; int bi = 0x7fffffff;
        for (int j = tid; j < m; j += nth) { double v = d[j]; if (v > bv || (v == bv && j < bi)) { bv = v; bi = j; } }
        // arg-max (largest value, smallest index among equals) on order-preserving integer keys with three warp REDUX
        // instructions per level instead of five rounds of 64-bit shuffles and fp64 compares; values <= 0 (exhausted
        // pivots carry -DBL_MAX) map to key 0 = "no pivot left".  One barrier: every warp reduces the warp results itself.
        warp_argmax_pos(bv, bi);
        if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_i[tid >> 5] = bi; }
        __syncthreads();
        bv = (tid & 31) < nwarps ? red_v[tid & 31] : 0.; bi = (tid & 31) < nwarps ? red_i[tid & 31] : 0x7fffffff;
        warp_argmax_pos(bv, bi);
        const double dp = bv; const int p = bi;
        if (k == 0) thresh = tol_rel * dp;
        if (!(dp > thresh) || !(dp > 0.)) { done = 1; break; }
        const int ks = min(k, srows);                // rows [0, ks) come from shared memory (a step then costs ~0.4 us instead of
                                                     // ~3.8 us of dependent L2 round trips), rows [ks, k) from the global factor
        if (k > ks) {
            for (int q = ks + tid; q < k; q += nth) lp[q] = At[(int64_t)q * ldm + p];
            __syncthreads();
        }
        const double inv = 1. / sqrt(dp);
        for (int j = tid; j < m; j += nth) {
            // even rows accumulate in s0, odd rows in s1, each in row order (the same sums whatever srows is)
            double s0 = 0., s1 = 0.;
            int q = 0;
            for (; q + 1 < ks; q += 2) {
                s0 += Ls[q * m + j] * Ls[q * m + p];
                s1 += Ls[(q + 1) * m + j] * Ls[(q + 1) * m + p];
            }
            if (q < ks) { s0 += Ls[q * m + j] * Ls[q * m + p]; ++q; }
            if (q < k && (q & 1)) { s1 += At[(int64_t)q * ldm + j] * lp[q]; ++q; }      // re-align to an even row
            // 8 independent loads in flight per pass (the factor lives in L2: one load at a time costs ~0.35 us)
            for (; q + 7 < k; q += 8) {
                double l[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) l[x] = At[(int64_t)(q + x) * ldm + j];
#pragma unroll
                for (int x = 0; x < 8; x += 2) { s0 += l[x] * lp[q + x]; s1 += l[x + 1] * lp[q + x + 1]; }
            }
            {
                double l[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) l[x] = (q + x < k) ? At[(int64_t)(q + x) * ldm + j] : 0.;
#pragma unroll
                for (int x = 0; x < 8; ++x) if (q + x < k) { if (x & 1) s1 += l[x] * lp[q + x]; else s0 += l[x] * lp[q + x]; }
            }
            const double gjp = G[(int64_t)p * m + j];      // G is symmetric: row p (coalesced) instead of column p
            double v = (gjp + (j == p ? shift : 0.) - (s0 + s1)) * inv;
            if (d[j] == -DBL_MAX) v = 0.;
            if (j == p) v = sqrt(dp);
            At[(int64_t)k * ldm + j] = v;
            if (k < srows) Ls[k * m + j] = v;
            d[j] = (j == p || d[j] == -DBL_MAX) ? -DBL_MAX : d[j] - v * v;
        }
        __syncthreads();
    }
    if (k == m) done = 1;
    if (d_out) for (int j = tid; j < m; j += nth) d_out[j] = d[j];      // state for the cooperative kernel to resume
    if (tid == 0) {
        r_out[blockIdx.x] = k; shift_out[blockIdx.x] = shift;
        if (thresh_out) *thresh_out = thresh;
        if (done_out) *done_out = done;
    }
}
}  // namespace

// Eigen-decomposition of `batch` PSD matrices G_z (m x m, contiguous).  Outputs per matrix z:
// factor vectors At_all + z*bsAt (cap_rows x ldm), sig_all / nrm_all + z*m, perm_all + z*cap_rows,
// r_dev[z] (device).  scratch_dbl needs batch * (cap_rows + 1) doubles.
int pbi_eig_psd_batched(pb_ctx* ctx, const double* G_all, int64_t m, int batch, double tol_rel, double shift_rel,
                        double* At_all, double* sig_all, double* nrm_all, int* perm_all, int* r_dev,
                        double* scratch_dbl, int* sweeps_out, int max_sweeps) {
    if (m <= 0 || m > 1024) PB_FAIL(ctx, PB_ERR_SHAPE, "eig_psd_batched: order %lld outside (0, 1024]", (long long)m);
    const int64_t ldm = (m + 7) / 8 * 8;
    const int64_t cap_rows = (m + JP - 1) / JP * JP;
    const int64_t bsAt = cap_rows * ldm;
    PB_CUDA(ctx, cudaMemsetAsync(At_all, 0, sizeof(double) * (size_t)bsAt * batch, ctx->stream));
    double* nrm2 = scratch_dbl; double* shift = scratch_dbl + (size_t)batch * cap_rows;
    {   // factor rows cached in shared memory: up to ~96 KB per CTA so that two matrices share an SM
        const int srows = (int)std::min<int64_t>(m, (96 * 1024) / (8 * m));
        const size_t sm = sizeof(double) * (size_t)m * (2 + srows);
        PB_TRY(pbi_smem_optin(ctx, (const void*)chol_small_kernel, sm));
        chol_small_kernel<<<batch, m > 256 ? 1024 : 256, sm, ctx->stream>>>(G_all, (int)m, At_all, ldm, bsAt, tol_rel, shift_rel, r_dev,
                                                                          shift, (int)m, nullptr, nullptr, nullptr, srows);
    }
    PB_LAUNCH_CHECK(ctx);
    const int nb = (int)(cap_rows / JB);
    const int nvec = (int)cap_rows;
    const double tol = std::max(4., sqrt((double)m)) * DBL_EPSILON;
    size_t jac_smem = sizeof(double) * 2 * JP * JPX;
    PB_TRY(pbi_smem_optin(ctx, (const void*)jacobi_pair_kernel, (size_t)(jac_smem)));
    int sweeps = 0; bool converged = false;
    const bool trust_single = ctx->eig_trust_single_pair && nb == 2;
    ctx->eig_trust_single_pair = false;
    for (; sweeps < max_sweeps && !converged; ++sweeps) {
        PB_CUDA(ctx, cudaMemsetAsync(ctx->d_int + 1, 0, sizeof(int), ctx->stream));
        const int rounds = nb > 2 ? nb - 1 : 1;
        for (int round = 0; round < rounds; ++round) {
            JacArgs ja{At_all, ldm, nb, round, tol, ctx->d_int + 1, g_inner_sweeps, bsAt, (g_cross_rounds && round > 0) ? 1 : 0, nullptr};
            jacobi_pair_kernel<<<dim3(nb / 2, batch), 256, jac_smem, ctx->stream>>>(ja);
            PB_LAUNCH_CHECK(ctx);
        }
        PB_CUDA(ctx, cudaMemcpyAsync(ctx->h_int + 1, ctx->d_int + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        converged = (ctx->h_int[1] == 0);
    }
    if (!converged && max_sweeps >= 30)
        PB_FAIL(ctx, PB_ERR_NOT_CONVERGED, "eig_psd_batched: Jacobi did not converge in %d sweeps", max_sweeps);
    norms_kernel<<<dim3(nvec, batch), 256, 0, ctx->stream>>>(At_all, ldm, nvec, nrm2, bsAt, cap_rows);
    PB_LAUNCH_CHECK(ctx);
    int npow2 = 1; while (npow2 < nvec) npow2 <<= 1;
    size_t sort_smem = (sizeof(double) + sizeof(int)) * (size_t)npow2;
    PB_TRY(pbi_smem_optin(ctx, (const void*)sort_kernel, (size_t)(sort_smem)));
    // r_eff is not needed for the batch: reuse the tail of the scratch as a dump (ints)
    int* r_eff = reinterpret_cast<int*>(shift + batch);
    sort_kernel<<<batch, 1024, sort_smem, ctx->stream>>>(nrm2, nvec, npow2, shift, sig_all, nrm_all, (int)m, perm_all, r_eff,
                                                        cap_rows, m, cap_rows);
    PB_LAUNCH_CHECK(ctx);
    if (sweeps_out) *sweeps_out = sweeps;
    return PB_OK;
}

int pbi_rank_batched(pb_ctx* ctx, const double* sig_all, int64_t m, int batch, double eps, int* nL_dev) {
    rank_kernel<<<batch, 32, 0, ctx->stream>>>(sig_all, (int)m, eps, nL_dev, m);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

// B_z (m x k, ldb = k) = factor columns / (nrm * (sig ? sig : 1)) for every matrix of the batch
int pbi_factor_to_right_batched(pb_ctx* ctx, const double* At_all, int64_t m, int k, int batch, const int* perm_all,
                                const double* nrm_all, const double* sig_all, double* B_all, int64_t bsB) {
    const int64_t ldm = (m + 7) / 8 * 8, cap_rows = (m + JP - 1) / JP * JP;
    int64_t n = m * k;
    factor_to_right_kernel<<<dim3((unsigned)((n + 255) / 256), batch), 256, 0, ctx->stream>>>(
        At_all, ldm, perm_all, nrm_all, sig_all, (int)m, k, B_all, k, cap_rows * ldm, m, cap_rows, bsB);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}


int pbi_chol_small_launch(pb_ctx* ctx, const double* G, int64_t m, double* At, int64_t ldm, double tol_rel, double shift_rel,
                          int* r_dev, double* shift_dev, int max_steps, double* d_out, double* thresh_out, int* done_out) {
    const int steps = std::min<int>(max_steps, (int)m);
    // one CTA: as many factor rows in shared memory as fit next to the 12 KB of static reduction buffers (m = 1000: 25 rows)
    const int srows = (int)std::max<int64_t>(0, std::min<int64_t>(steps, (210 * 1024 - 16 * m) / (8 * m)));
    const size_t sm = sizeof(double) * (size_t)m * (2 + srows);
    PB_TRY(pbi_smem_optin(ctx, (const void*)chol_small_kernel, sm));
    chol_small_kernel<<<1, m > 256 ? 1024 : 256, sm, ctx->stream>>>(G, (int)m, At, ldm, 0, tol_rel, shift_rel, r_dev, shift_dev, steps, d_out,
                                                                   thresh_out, done_out, srows);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}
