// TSQR accuracy mode of the POD (BASELINE.json north_star: "a TSQR accuracy mode for ill-conditioned snapshot
// sets"; SURVEY.md section 8e: local QR per row shard, all-gather of the R factors, QR of the stack).
//
//   A (rows x m, tall)  --blocked Householder QR, R only-->  R1 (m x m)          [qr_panel_kernel + DMMA GEMMs]
//   R1                  --Householder QR with column pivoting, rank revealing-->  r factor vectors  [qrcp_kernel]
//   one-sided block Jacobi on the r vectors (eig.cu)  -->  sigma, right singular vectors Z;  V = A Z / sigma
//
// Working on A itself (never on A^T A) keeps the singular values and the right singular vectors at the accuracy of
// LAPACK's dgesdd, which is what the reference calls (poduqnn/pod.py:16); V = A z / sigma is the reference's own
// formula for the basis (pod.py:42-45), so Q is never formed or stored.
//
// Big QR: right-looking, outer panels of QNB = 64 columns on a working copy W of A.
//   panel (cooperative kernel, CTAs own row ranges): inner panels of 16 columns, column by column with ONE grid
//     barrier per column -- every CTA sends the partial dot products of the current column with the columns to its
//     right (rows below the diagonal row) and the owner of the diagonal row sends that row; with those 2 x 16 numbers
//     every CTA forms the same reflector (LAPACK dlarfg formulas) and updates its rows, accumulating the partial dot
//     products of the NEXT column in the same pass.  The inner block reflector is applied to the rest of the outer
//     panel on DMMA (Z = Y^T C reduced over the grid, X from tau and striu(Y^T Y) by forward substitution --
//     T^-1 = diag(1/tau) + striu(Y^T Y), so T is never formed -- then C -= Y X).
//   trailing update: Z = Y^T C and C -= Y X with the tall-skinny DMMA kernels of gemm.cu.
// Small QRCP: CTAs own groups of 8 columns; a step picks the remaining column of largest norm (exactly recomputed in
//   the previous update pass), reflects it and updates the remaining columns -- columns are never moved, so the rows
//   of the result are the factor vectors in the ORIGINAL column order; one grid barrier per step; stops at the
//   numerical rank.
#include "common.cuh"
#include "tile_ops.cuh"
#include <cooperative_groups.h>
#include <cfloat>
#include <cstdlib>
#include <algorithm>
namespace cg = cooperative_groups;

namespace {
using pbt::dmma884;

constexpr int QNB = 64;     // outer panel width (columns)
constexpr int QIB = 16;     // inner panel width
constexpr int QT = 256;     // threads per CTA
constexpr int ZE = QIB * QNB;   // entries of the inner Z block: 16 x (16 + up to 48)

// explicit entry of the unit-lower-trapezoidal reflector block whose first column is c0: column a has its diagonal
// in row c0 + a
__device__ __forceinline__ double yval(const double* __restrict__ W, int64_t ldw, int64_t r, int c0, int a) {
    const int64_t d = (int64_t)c0 + a;
    return r > d ? W[r * ldw + d] : (r == d ? 1. : 0.);
}

// sum of v[0..16) over the CTA, result in out[0..16) (shared); fixed order.  Warp stage: a butterfly that halves the
// number of values a lane carries at every step (8 + 4 + 2 + 1 + 1 = 16 exchanges instead of 16 x 5): after it lane l
// holds the warp sum of value (l >> 1) & 15 ... see the index algebra below.
__device__ __forceinline__ void block_sum16(double (&v)[QIB], double (*red)[QIB], double* out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // stage with xor 16: lanes with bit 4 clear keep values 0..7, the others 8..15
    double a8[8];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double send = up ? v[k] : v[k + 8];
            const double keep = up ? v[k + 8] : v[k];
            a8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    double a4[4];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double send = up ? a8[k] : a8[k + 4];
            const double keep = up ? a8[k + 4] : a8[k];
            a4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    double a2[2];
    {
        const bool up = lane & 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double send = up ? a4[k] : a4[k + 2];
            const double keep = up ? a4[k + 2] : a4[k];
            a2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    double a1;
    {
        const bool up = lane & 2;
        const double send = up ? a2[0] : a2[1];
        const double keep = up ? a2[1] : a2[0];
        a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    // lane l now holds the warp sum of value index 8 b4 + 4 b3 + 2 b2 + b1 (b_i = bit i of l)
    if ((lane & 1) == 0) red[warp][((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = a1;
    __syncthreads();
    if (tid < QIB) {
        double s = 0.;
#pragma unroll
        for (int w = 0; w < QT / 32; ++w) s += red[w][tid];
        out[tid] = s;
    }
    __syncthreads();
}

struct QrPanelArgs {
    double* W; int64_t ldw; int64_t rows;    // working copy (rows >= padded column count, zero padded)
    int J; int nbo;                          // outer panel: columns [J, J + nbo), nbo a multiple of 16, <= 64
    double* tau;                             // tau + J
    double* part;                            // [2][G][16] partial dot products, double buffered
    double* diag;                            // [2][16] the published diagonal row
    double* zpart;                           // [G][ZE] per-CTA partial Z of the inner apply
    double* zfin;                            // [ZE]
    int use_smem;                            // the CTA's rows of an inner panel fit shared memory: the 16 column steps
                                             // run on the staged copy (pitch QP), written back once
};
constexpr int QP = QIB + 2;                  // shared-memory row pitch: 16-byte aligned, conflict-free double2 access

// Row pass of column step C of an inner panel (C is a compile-time constant: no predicate / select chains in the loop,
// only the 16-byte chunks that hold columns >= C are touched).  Applies the reflector to my rows, stores v in column C,
// accumulates the partial dot products of column C + 1 with the columns to its right (rows below the next diagonal row)
// and lets the owner of the next diagonal row publish it.
template <int C>
__device__ __forceinline__ void qr_row_pass(double* __restrict__ W, int64_t ldw, double* Ps, int use_smem, int c0, int64_t lo,
                                            int64_t hi, int64_t d, double beta, double tau, double scale,
                                            const double* __restrict__ gsh, const double* __restrict__ dsh,
                                            double* __restrict__ diag_next, double (&g)[QIB], int tid) {
    constexpr int K0 = C / 2;                       // first 16-byte chunk that holds a column >= C
    double w[QIB];
#pragma unroll
    for (int k = 0; k < QIB; ++k) w[k] = k > C ? dsh[k] + gsh[k] * scale : 0.;     // v^T (column k)
    const int64_t rb = lo > d ? lo : d;
    for (int64_t r = rb + tid; r < hi; r += QT) {
        double x[QIB];
        double2* p = use_smem ? reinterpret_cast<double2*>(Ps + (size_t)(r - lo) * QP) : reinterpret_cast<double2*>(W + r * ldw + c0);
#pragma unroll
        for (int k = K0; k < QIB / 2; ++k) { double2 v = p[k]; x[2 * k] = v.x; x[2 * k + 1] = v.y; }
        if (r == d) {
            x[C] = beta;
#pragma unroll
            for (int k = C + 1; k < QIB; ++k) x[k] -= tau * w[k];
        } else {
            const double v = x[C] * scale, tv = tau * v;
            x[C] = v;
#pragma unroll
            for (int k = C + 1; k < QIB; ++k) x[k] -= tv * w[k];
            if (C + 1 < QIB) {
                if (r == d + 1) {
#pragma unroll
                    for (int k = C + 1; k < QIB; ++k) diag_next[k] = x[k];
                } else {
                    const double xn = x[C + 1 < QIB ? C + 1 : C];
#pragma unroll
                    for (int k = C + 1; k < QIB; ++k) g[k] += xn * x[k];
                }
            }
        }
#pragma unroll
        for (int k = K0; k < QIB / 2; ++k) p[k] = make_double2(x[2 * k], x[2 * k + 1]);
    }
}

__global__ void __launch_bounds__(QT, 1) qr_panel_kernel(QrPanelArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double Ps[];            // the staged inner panel (use_smem), pitch QP
    __shared__ double red[QT / 32][QIB];
    __shared__ double red2[QIB][QIB];
    __shared__ double gsh[QIB], dsh[QIB], tau_sh[QIB], outsh[QIB];
    __shared__ double Zs[ZE];
    __shared__ double Xs[QIB][QNB - QIB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lk = lane & 3, lg = lane >> 2;
    const int G = gridDim.x, b = blockIdx.x;
    double* __restrict__ W = a.W; const int64_t ldw = a.ldw;
    // row range of this CTA over the active rows [J, rows)
    int64_t rpc = (a.rows - a.J + G - 1) / G; rpc = (rpc + 7) / 8 * 8;
    const int64_t lo = (int64_t)a.J + (int64_t)b * rpc;
    const int64_t hi = lo < a.rows ? (lo + rpc < a.rows ? lo + rpc : a.rows) : lo;
    int par = 0;
#ifdef PB_QR_TIMING
    long long tq[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t0q, t1q;
#define TQ(i) do { t1q = clock64(); tq[i] += t1q - t0q; t0q = t1q; } while (0)
    t0q = clock64();
#else
#define TQ(i) do { } while (0)
#endif

    for (int i0 = 0; i0 < a.nbo; i0 += QIB) {
        const int c0 = a.J + i0;
        // ---- prime: partial dots of column c0 with the 16 columns, rows below the diagonal row c0
        {
            double g[QIB];
#pragma unroll
            for (int k = 0; k < QIB; ++k) g[k] = 0.;
            const int64_t rb = lo > c0 ? lo : (int64_t)c0;
            for (int64_t r = rb + tid; r < hi; r += QT) {
                double x[QIB];
                const double2* src = reinterpret_cast<const double2*>(W + r * ldw + c0);
#pragma unroll
                for (int k = 0; k < QIB / 2; ++k) { double2 v = src[k]; x[2 * k] = v.x; x[2 * k + 1] = v.y; }
                if (a.use_smem) {
                    double2* dst = reinterpret_cast<double2*>(Ps + (size_t)(r - lo) * QP);
#pragma unroll
                    for (int k = 0; k < QIB / 2; ++k) dst[k] = make_double2(x[2 * k], x[2 * k + 1]);
                }
                if (r == c0) {
#pragma unroll
                    for (int k = 0; k < QIB; ++k) a.diag[par * QIB + k] = x[k];
                } else {
#pragma unroll
                    for (int k = 0; k < QIB; ++k) g[k] += x[0] * x[k];
                }
            }
            block_sum16(g, red, outsh);
            if (tid < QIB) a.part[((size_t)par * G + b) * QIB + tid] = outsh[tid];
        }
        TQ(0);
        grid.sync();
        TQ(1);

        for (int c = 0; c < QIB; ++c) {
            const int64_t d = (int64_t)c0 + c;
            // ---- every CTA reduces the partials in the same order
            {
                const int k = tid & 15, grp = tid >> 4;
                double s = 0.;
                for (int bb = grp; bb < G; bb += QT / 16) s += __ldcg(a.part + ((size_t)par * G + bb) * QIB + k);
                red2[grp][k] = s;
                __syncthreads();
                if (tid < QIB) {
                    double t = 0.;
#pragma unroll
                    for (int q = 0; q < QT / 16; ++q) t += red2[q][tid];
                    gsh[tid] = t; dsh[tid] = __ldcg(a.diag + par * QIB + tid);
                }
                __syncthreads();
            }
            TQ(2);
            // ---- reflector (LAPACK dlarfg): beta = -sign(alpha) hypot(alpha, xnorm), tau = (beta - alpha) / beta
            const double alpha = dsh[c], xn2 = gsh[c];
            double beta = alpha, tau = 0., scale = 0.;
            if (xn2 > 0.) {
                beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
                tau = (beta - alpha) / beta;
                scale = 1. / (alpha - beta);
            }
            if (tid == 0) { tau_sh[c] = tau; if (b == 0) a.tau[i0 + c] = tau; }
            // ---- update my rows; accumulate the partials of column c + 1
            double g[QIB];
#pragma unroll
            for (int k = 0; k < QIB; ++k) g[k] = 0.;
            double* dn = a.diag + (par ^ 1) * QIB;
#define PB_QR_CASE(CC) case CC: qr_row_pass<CC>(W, ldw, Ps, a.use_smem, c0, lo, hi, d, beta, tau, scale, gsh, dsh, dn, g, tid); break;
            switch (c) {
                PB_QR_CASE(0) PB_QR_CASE(1) PB_QR_CASE(2) PB_QR_CASE(3) PB_QR_CASE(4) PB_QR_CASE(5) PB_QR_CASE(6) PB_QR_CASE(7)
                PB_QR_CASE(8) PB_QR_CASE(9) PB_QR_CASE(10) PB_QR_CASE(11) PB_QR_CASE(12) PB_QR_CASE(13) PB_QR_CASE(14)
                default: qr_row_pass<15>(W, ldw, Ps, a.use_smem, c0, lo, hi, d, beta, tau, scale, gsh, dsh, dn, g, tid); break;
            }
#undef PB_QR_CASE
            TQ(3);
            if (c + 1 < QIB) {
                block_sum16(g, red, outsh);
                if (tid < QIB) a.part[((size_t)(par ^ 1) * G + b) * QIB + tid] = outsh[tid];
                par ^= 1;
                TQ(4);
                grid.sync();        // (a flag-per-CTA exchange instead of this barrier measured slower: 1.14 vs 0.83 ms per panel)
                TQ(1);
            }
        }
        par ^= 1;
        __syncthreads();
        if (a.use_smem) {                                           // staged inner panel back to the working copy
            const int64_t rb = lo > c0 ? lo : (int64_t)c0;
            for (int64_t r = rb + tid; r < hi; r += QT) {
                const double2* src = reinterpret_cast<const double2*>(Ps + (size_t)(r - lo) * QP);
                double2* dst = reinterpret_cast<double2*>(W + r * ldw + c0);
#pragma unroll
                for (int k = 0; k < QIB / 2; ++k) dst[k] = src[k];
            }
            __syncthreads();
        }

        TQ(9);
        // ---- apply the inner block reflector to the rest of the outer panel
        const int ncols = a.nbo - i0 - QIB;
        if (ncols > 0) {                                           // uniform
            const int c1 = c0 + QIB;
            const int nblk = (QIB + ncols) / 8;
            const int64_t lo2 = lo > c0 ? lo : (int64_t)c0;
            // entry (r, col) of the explicit reflector block: from the staged copy when there is one
            auto yv = [&](int64_t r, int col) -> double {
                const int64_t d = (int64_t)c0 + col;
                if (r > d) return a.use_smem ? Ps[(size_t)(r - lo) * QP + col] : W[r * ldw + d];
                return r == d ? 1. : 0.;
            };
            // Z = Y^T [Y | C] over my rows: two 4-row steps per warp iteration, all loads issued before the DMMAs
            double zacc[2][QNB / 8][2];
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int u = 0; u < QNB / 8; ++u) { zacc[t][u][0] = 0.; zacc[t][u][1] = 0.; }
            for (int64_t r0 = lo2 + 4 * warp; r0 < hi; r0 += 8 * (QT / 32)) {
                double a0[2], a1[2], bv[2][QNB / 8 - 2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t r = r0 + h * 4 * (QT / 32) + lk;
                    const bool ok = r < hi;
                    a0[h] = ok ? yv(r, lg) : 0.;
                    a1[h] = ok ? yv(r, 8 + lg) : 0.;
#pragma unroll
                    for (int u = 2; u < QNB / 8; ++u)
                        bv[h][u - 2] = (ok && u < nblk) ? W[r * ldw + c1 + (u - 2) * 8 + lg] : 0.;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int u = 0; u < QNB / 8; ++u) {
                        if (u < nblk) {
                            const double b_ = u == 0 ? a0[h] : (u == 1 ? a1[h] : bv[h][u >= 2 ? u - 2 : 0]);
                            dmma884(zacc[0][u][0], zacc[0][u][1], a0[h], b_);
                            dmma884(zacc[1][u][0], zacc[1][u][1], a1[h], b_);
                        }
                    }
            }
            TQ(6);
            // CTA sum in warp order (fixed), straight into Zs
            for (int e = tid; e < ZE; e += QT) Zs[e] = 0.;
            __syncthreads();
            for (int wq = 0; wq < QT / 32; ++wq) {
                if (warp == wq) {
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int u = 0; u < QNB / 8; ++u) {
                            Zs[(t * 8 + lg) * QNB + u * 8 + 2 * lk] += zacc[t][u][0];
                            Zs[(t * 8 + lg) * QNB + u * 8 + 2 * lk + 1] += zacc[t][u][1];
                        }
                }
                __syncthreads();
            }
            for (int e = tid; e < ZE; e += QT) a.zpart[(size_t)b * ZE + e] = Zs[e];
            grid.sync();
            {   // second stage: CTA b finishes entries [b * epc, (b + 1) * epc)
                const int epc = (ZE + G - 1) / G;
                for (int el = warp; el < epc; el += QT / 32) {
                    const int e = b * epc + el;
                    if (e < ZE) {
                        double s = 0.;
                        for (int bb = lane; bb < G; bb += 32) s += __ldcg(a.zpart + (size_t)bb * ZE + e);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        if (lane == 0) a.zfin[e] = s;
                    }
                }
            }
            grid.sync();
            for (int e = tid; e < ZE; e += QT) Zs[e] = __ldcg(a.zfin + e);
            __syncthreads();
            TQ(7);
            // X = T^T Z: X_i = tau_i (Z_i - sum_{k<i} S_ki X_k), S = striu(Y^T Y) = Zs[:, 0:16]
            if (tid < ncols) {
                for (int i = 0; i < QIB; ++i) {
                    double s = Zs[i * QNB + QIB + tid];
                    for (int k = 0; k < i; ++k) s -= Zs[k * QNB + i] * Xs[k][tid];
                    Xs[i][tid] = tau_sh[i] * s;
                }
            }
            __syncthreads();
            TQ(8);
            // C -= Y X on my rows: two 8-row groups per warp iteration
            for (int64_t r0 = lo2 + 8 * warp; r0 < hi; r0 += 16 * (QT / 32)) {
                double af[2][4];
                double2 cc[2][(QNB - QIB) / 8];
                bool ok[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t rr = r0 + h * 8 * (QT / 32) + lg;
                    ok[h] = rr < hi;
#pragma unroll
                    for (int q = 0; q < 4; ++q) af[h][q] = ok[h] ? yv(rr, 4 * q + lk) : 0.;
#pragma unroll
                    for (int u = 0; u < (QNB - QIB) / 8; ++u)
                        cc[h][u] = (ok[h] && u * 8 < ncols) ? *reinterpret_cast<const double2*>(W + rr * ldw + c1 + u * 8 + 2 * lk)
                                                            : make_double2(0., 0.);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t rr = r0 + h * 8 * (QT / 32) + lg;
#pragma unroll
                    for (int u = 0; u < (QNB - QIB) / 8; ++u) {
                        if (u * 8 < ncols) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) dmma884(cc[h][u].x, cc[h][u].y, af[h][q], -Xs[4 * q + lk][u * 8 + lg]);
                            if (ok[h]) *reinterpret_cast<double2*>(W + rr * ldw + c1 + u * 8 + 2 * lk) = cc[h][u];
                        }
                    }
                }
            }
            __syncthreads();
        }
        TQ(5);
    }
#ifdef PB_QR_TIMING
    if (tid == 0 && (b == 0 || b == G / 2) && a.J == 256)
        printf("cta %d J %d: prime %lld  gridsync %lld  gather %lld  rowpass %lld  blocksum+publish %lld  writeback %lld | apply: zpass %lld  reduce+2syncs %lld  solve %lld  update %lld (cycles)\n", b, a.J,
               tq[0], tq[1], tq[2], tq[3], tq[4], tq[9], tq[6], tq[7], tq[8], tq[5]);
#endif
}

// After a panel is factored its columns of W hold R (on and above the diagonal, rows 0 .. J + a) and the reflector
// vectors (below).  Move the R entries to the output and put the explicit unit-lower-trapezoidal form in their place
// (zeros above, ones on the diagonal): the panel columns of W then ARE the block Y, so that [Y^T Y | Y^T C] is one TN
// product over W[:, J:] and the update reads Y in place -- no copy of Y, no separate Y^T Y launch.
__global__ void qr_save_r_make_y_kernel(double* __restrict__ W, int64_t ldw, int J, int nbo, int64_t m,
                                        double* __restrict__ R, int64_t ldr, const int* __restrict__ colperm) {
    // colperm (rank-revealing mode): position d of the working copy holds column colperm[d] of the input
    const int64_t total = (int64_t)(J + nbo) * nbo;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / nbo; const int a = (int)(idx % nbo);
        const int64_t d = (int64_t)J + a;
        if (r > d) continue;
        double* p = W + r * ldw + d;
        const int64_t dc = colperm ? colperm[d] : d;
        if (r < m && dc < m) R[r * ldr + dc] = *p;
        *p = (r == d) ? 1. : 0.;
    }
}

// T^T (nbo x nbo lower triangular, ld 64) from tau and S = Y^T Y: the recurrence X_i = tau_i (Z_i - sum_{k<i} S_ki X_k)
// applied to Z = I (T^-1 = diag(1/tau) + striu(S); tau_i = 0 gives a zero row, i.e. H_i = I).  One CTA; 4 threads share the
// k-sum of a column; S is staged in shared memory (its loads were the critical path of the 64 serial steps).
__global__ void __launch_bounds__(4 * QNB) qr_tt_kernel(const double* __restrict__ S, int64_t lds, int nbo,
                                                        const double* __restrict__ tau, double* __restrict__ Tt) {
    extern __shared__ __align__(16) double tts[];
    double (*Xs)[QNB + 1] = reinterpret_cast<double (*)[QNB + 1]>(tts);
    double (*Ss)[QNB + 1] = reinterpret_cast<double (*)[QNB + 1]>(tts + QNB * (QNB + 1));
    __shared__ double ts[QNB];
    const int tid = threadIdx.x, j = tid >> 2, q = tid & 3;
    for (int e = tid; e < QNB * QNB; e += 4 * QNB) {
        const int i = e / QNB, c = e % QNB;
        Xs[i][c] = 0.;
        Ss[c][i] = (i < nbo && c < nbo) ? S[(int64_t)i * lds + c] : 0.;        // transposed: Ss[i][k] = S_ki
    }
    if (tid < QNB) ts[tid] = tid < nbo ? tau[tid] : 0.;
    __syncthreads();
    for (int i = 0; i < nbo; ++i) {
        double s = 0.;
        for (int k = j + q; k < i; k += 4) s -= Ss[i][k] * Xs[k][j];             // X_kj = 0 for k < j
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0 && j < nbo && i >= j) Xs[i][j] = ts[i] * ((i == j ? 1. : 0.) + s);
        __syncwarp();
    }
    __syncthreads();
    for (int e = tid; e < QNB * QNB; e += 4 * QNB) {
        const int i = e / QNB, c = e % QNB;
        Tt[e] = (i < nbo && c < nbo) ? Xs[i][c] : 0.;
    }
}

// -X = -T^T Z in place for a tile of QAT trailing columns per CTA (narrow tiles: the product is tiny, the point is to
// spread it over enough CTAs)
constexpr int QAT = 16;
__global__ void __launch_bounds__(QT) qr_applyt_kernel(double* __restrict__ Z, int64_t ldz, int64_t ncols,
                                                       const double* __restrict__ Tt, int nbo) {
    __shared__ double Zs[QNB][QAT + 1];
    __shared__ double Ts[QNB][QNB + 1];
    const int tid = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.x * QAT;
    for (int e = tid; e < QNB * QAT; e += QT) {
        const int k = e / QAT, jj = e % QAT;
        Zs[k][jj] = (k < nbo && j0 + jj < ncols) ? Z[(int64_t)k * ldz + j0 + jj] : 0.;
    }
    for (int e = tid; e < QNB * QNB; e += QT) Ts[e / QNB][e % QNB] = Tt[e];
    __syncthreads();
    const int jj = tid % QAT, part = tid / QAT;               // 16 row groups
    if (j0 + jj >= ncols) return;
    for (int i = part; i < nbo; i += QT / QAT) {
        double s0 = 0., s1 = 0.;
        int k = 0;
        for (; k + 1 <= i; k += 2) { s0 += Ts[i][k] * Zs[k][jj]; s1 += Ts[i][k + 1] * Zs[k + 1][jj]; }
        if (k <= i) s0 += Ts[i][k] * Zs[k][jj];
        Z[(int64_t)i * ldz + j0 + jj] = -(s0 + s1);            // -X: the update kernel accumulates C += Y (-X)
    }
}

// W (rowsW x ldw) = A (rows x m) zero padded
__global__ void qr_copy_pad_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int64_t m,
                                   double* __restrict__ W, int64_t ldw, int64_t rowsW) {
    const int64_t total = rowsW * ldw;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / ldw, c = idx % ldw;
        W[idx] = (r < rows && c < m) ? A[r * lda + c] : 0.;
    }
}

// the same with the columns gathered through colmap (position -> source column): the first block pivot of the
// rank-revealing mode is applied while the working copy is made instead of by a swap pass over it
__global__ void qr_copy_pad_perm_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int64_t m,
                                        double* __restrict__ W, int64_t ldw, int64_t rowsW, const int* __restrict__ colmap) {
    const int64_t total = rowsW * ldw;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / ldw; const int c = (int)(idx % ldw);
        const int src = colmap[c];
        W[idx] = (r < rows && src < m) ? A[r * lda + src] : 0.;
    }
}

// R (m x m, ld ldr) = upper triangle of W, zero below; rows [m, rows_out) of the destination are zeroed as well
__global__ void qr_extract_r_kernel(const double* __restrict__ W, int64_t ldw, int64_t m, double* __restrict__ R,
                                    int64_t ldr, int64_t rows_out) {
    const int64_t total = rows_out * ldr;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / ldr, j = idx % ldr;
        R[idx] = (i < m && j < m && j >= i) ? W[i * ldw + j] : 0.;
    }
}

// ---------------------------------------------------------------------------
// Rank-revealing mode of the big QR: block column pivoting + early termination.
//
// POD snapshot matrices have low numerical rank (C2: 36 of 1000 columns at m eps, C5: ~190 of 4096), and the modes
// need only a factor F with F^T F = A^T A to working accuracy.  After k panels the working copy holds
// Q^T A P = [R11 R12; 0 S]: once every column of the residual S is below tol * (largest column norm of A), S is at
// the level of the QR's own rounding error and is dropped -- F = [R11 R12] P^T has 64 k rows instead of m, and the
// remaining (m/64 - k) panels (almost all of the 2 n m^2 flops) are never run.  Before each panel the columns with
// the largest residual norms are swapped into the panel position (block pivoting: it only has to expose the rank
// early, any column order gives a valid QR), so an unlucky column order cannot delay the stop.  Residual norms are
// downdated with the new R rows after each panel (LAPACK dgeqp3's scheme); downdated squares cannot resolve anything
// below eps * the original, so the stop decision is taken only on exactly recomputed norms (one read pass over the
// trailing matrix, run when the downdated maximum falls below 1e-6 of the start).  A full-rank matrix never triggers
// the recomputation and costs the same as the plain blocked QR.
// ---------------------------------------------------------------------------
constexpr int CN_COLS = 256;        // columns per CTA of the norm kernel (128 threads x double2)

// part[rb][c] = sum over the rows of block rb (within [r0, rows)) of W[r][c0 + c]^2
__global__ void __launch_bounds__(128) qr_colnorm_part_kernel(const double* __restrict__ W, int64_t ldw, int64_t r0,
                                                              int64_t rows, int c0, int ncols, double* __restrict__ part,
                                                              int64_t ldp) {
    const int c = blockIdx.x * CN_COLS + 2 * threadIdx.x;
    if (c >= ncols) return;
    const int nrb = gridDim.y, rb = blockIdx.y;
    const int64_t per = (rows - r0 + nrb - 1) / nrb;
    const int64_t lo = r0 + (int64_t)rb * per, hi = lo + per < rows ? lo + per : rows;
    const double* p = W + (int64_t)c0 + c;
    double s0 = 0., s1 = 0.;
    int64_t r = lo;
    for (; r + 4 <= hi; r += 4) {
        const double2 v0 = *reinterpret_cast<const double2*>(p + r * ldw);
        const double2 v1 = *reinterpret_cast<const double2*>(p + (r + 1) * ldw);
        const double2 v2 = *reinterpret_cast<const double2*>(p + (r + 2) * ldw);
        const double2 v3 = *reinterpret_cast<const double2*>(p + (r + 3) * ldw);
        s0 += v0.x * v0.x; s1 += v0.y * v0.y; s0 += v1.x * v1.x; s1 += v1.y * v1.y;
        s0 += v2.x * v2.x; s1 += v2.y * v2.y; s0 += v3.x * v3.x; s1 += v3.y * v3.y;
    }
    for (; r < hi; ++r) { const double2 v = *reinterpret_cast<const double2*>(p + r * ldw); s0 += v.x * v.x; s1 += v.y * v.y; }
    part[(int64_t)rb * ldp + c] = s0; part[(int64_t)rb * ldp + c + 1] = s1;
}
__global__ void qr_colnorm_reduce_kernel(const double* __restrict__ part, int nrb, int64_t ldp, int ncols, double* __restrict__ cn2) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double s = 0.;
    for (int rb = 0; rb < nrb; ++rb) s += part[(int64_t)rb * ldp + c];       // fixed order
    cn2[c] = s;
}

// cn2[j] -= sum_{k in [J, J + nbo)} W[k][j]^2 for the trailing columns j >= J + nbo (rows J .. J + nbo - 1 are final R rows)
__global__ void qr_downdate_kernel(const double* __restrict__ W, int64_t ldw, int J, int nbo, int mp, double* __restrict__ cn2) {
    const int j = J + nbo + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= mp) return;
    double s = 0.;
    for (int k = 0; k < nbo; ++k) { const double v = W[(int64_t)(J + k) * ldw + j]; s += v * v; }
    const double d = cn2[j] - s;
    cn2[j] = d > 0. ? d : 0.;
}

struct QrPivArgs {
    double* cn2; int* colperm; int J, nbo, mp;
    double tol2;            // stop when the largest exact residual norm^2 <= tol2 * first
    double* first;          // device scalar: largest column norm^2 of the input (written at J == 0)
    int exact;              // cn2[J..) was recomputed exactly just now
    int pivot;              // block pivoting on / off
    int* out;               // [0] stop, [1] exact norms wanted, [2] number of swaps; [4 + 2 s], [5 + 2 s]: positions of swap s
};
constexpr int PIVT = 1024;

// One CTA.  Decides stop / recompute / the column swaps of the next panel.
// Choice of the nbo panel columns: the candidates are the trailing columns whose residual norm is within 4x of the
// largest one; nbo of them are taken EVENLY SPREAD over their positions.  Taking simply the nbo largest norms fails on
// smooth snapshot sets (neighbouring columns -- consecutive time steps, neighbouring DCT sample points -- have
// neighbouring norms and are nearly parallel, so a panel made of them exposes few new directions: measured 27 panels
// instead of 4 on the 250 000 x 4096 DCT matrix).  Fewer than nbo candidates: the nbo largest norms.  Columns that are
// already in the panel window and wanted stay; the others trade places with the wanted columns outside, in order.
__global__ void __launch_bounds__(PIVT) qr_pivot_kernel(QrPivArgs a) {
    extern __shared__ __align__(16) double pv[];        // [n] residual norms^2 of the trailing columns
    __shared__ double redv[PIVT / 32];
    __shared__ int scan[PIVT / 32];
    __shared__ int wlist[QNB], cand[QNB];
    __shared__ double t_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.mp - a.J, nbo = a.nbo;
    double mx = 0.;
    for (int j = tid; j < n; j += PIVT) { const double v = a.cn2[a.J + j]; pv[j] = v; mx = v > mx ? v : mx; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, mx, o); mx = t > mx ? t : mx; }
    if (lane == 0) redv[warp] = mx;
    __syncthreads();
    mx = 0.;
    for (int w = 0; w < PIVT / 32; ++w) mx = redv[w] > mx ? redv[w] : mx;
    double first = mx;
    if (a.J == 0) { if (tid == 0) *a.first = mx; } else first = *a.first;
    if (tid == 0) { a.out[0] = 0; a.out[1] = 0; a.out[2] = 0; }
    if (a.exact) { if (!(mx > a.tol2 * first)) { if (tid == 0) a.out[0] = 1; return; } }      // (also a zero matrix)
    else if (!(mx >= 1e-9 * first)) { if (tid == 0) a.out[1] = 1; return; }              // too small for downdated values
    if (!a.pivot || n <= nbo) return;
    if (tid < QNB) { wlist[tid] = -1; cand[tid] = -1; }
    // ---- candidates: within 4x of the largest norm (16x in the squares) -- widened by further factors of 4 while there
    // are fewer candidates than panel columns (after a panel the residual is largest BETWEEN the columns just taken, in
    // narrow clusters; a panel of cluster neighbours would again expose few directions); ordinal of each by position
    const int per = (n + PIVT - 1) / PIVT, j0 = tid * per, j1 = j0 + per < n ? j0 + per : n;
    double thr = mx;
    int cnt = 0, inc = 0, base = 0, nS = 0;
    for (int widen = 0; widen < 8 && nS < nbo; ++widen) {
        thr *= 0.0625;
        cnt = 0;
        for (int j = j0; j < j1; ++j) cnt += pv[j] >= thr ? 1 : 0;
        inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __syncthreads();
        if (lane == 31) scan[warp] = inc;
        __syncthreads();
        base = 0; nS = 0;
        for (int w = 0; w < PIVT / 32; ++w) { if (w < warp) base += scan[w]; nS += scan[w]; }
    }
    if (nS >= nbo) {
        int q = base + inc - cnt;                       // ordinal of my first candidate
        for (int j = j0; j < j1; ++j) {
            if (pv[j] >= thr) {
                const int w0 = (int)(((long long)q * nbo) / nS), w1 = (int)(((long long)(q + 1) * nbo) / nS);
                if (w1 != w0) wlist[w0] = j;            // the w0-th wanted column
                ++q;
            }
        }
    } else {
        // ---- fewer candidates than panel columns: the nbo largest norms (rank by counting; ties by position)
        for (int j = tid; j < n; j += PIVT) {
            const double vj = pv[j];
            int ra = 0;
            for (int k = 0; k < n; ++k) { const double vk = pv[k]; ra += (vk > vj || (vk == vj && k < j)) ? 1 : 0; }
            if (ra < nbo) wlist[ra] = j;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // window positions that are not wanted <-> wanted columns outside the window, both in ascending order
        bool inwin[QNB];
        for (int w = 0; w < nbo; ++w) inwin[w] = false;
        int nc = 0;
        // wlist is ascending in the spread branch; in the rank branch sort the outside ones by position (<= 64 entries)
        for (int i = 0; i < nbo; ++i) {
            const int j = wlist[i];
            if (j < 0) continue;
            if (j < nbo) inwin[j] = true;
            else { int k = nc++; while (k > 0 && cand[k - 1] > j) { cand[k] = cand[k - 1]; --k; } cand[k] = j; }
        }
        int ns = 0, ci = 0;
        for (int w = 0; w < nbo && ci < nc; ++w) {
            if (inwin[w]) continue;
            const int c = cand[ci++];
            a.out[4 + 2 * ns] = a.J + w; a.out[5 + 2 * ns] = a.J + c; ++ns;
            const double tv = a.cn2[a.J + w]; a.cn2[a.J + w] = a.cn2[a.J + c]; a.cn2[a.J + c] = tv;
            const int tp = a.colperm[a.J + w]; a.colperm[a.J + w] = a.colperm[a.J + c]; a.colperm[a.J + c] = tp;
        }
        a.out[2] = ns;
    }
}

// W[r][a_s] <-> W[r][b_s] for every row and every swap of the list (disjoint pairs)
__global__ void qr_swap_cols_kernel(double* __restrict__ W, int64_t ldw, int64_t rows, const int* __restrict__ sw, int ns) {
    const int64_t total = rows * ns;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / ns; const int s = (int)(idx % ns);
        double* pa = W + r * ldw + sw[2 * s]; double* pb = W + r * ldw + sw[2 * s + 1];
        const double t = *pa; *pa = *pb; *pb = t;
    }
}

__global__ void qr_iota_kernel(int* __restrict__ p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// rows [0, K) of the columns that were never factored (positions >= K): R12, back in the input's column order
__global__ void qr_gather_tail_kernel(const double* __restrict__ W, int64_t ldw, int K, int mp, int64_t m,
                                      const int* __restrict__ colperm, double* __restrict__ R, int64_t ldr) {
    const int nt = mp - K;
    const int64_t total = (int64_t)K * nt;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / nt; const int j = K + (int)(idx % nt);
        const int64_t jc = colperm ? colperm[j] : j;
        if (i < m && jc < m) R[i * ldr + jc] = W[i * ldw + j];
    }
}

// ---------------------------------------------------------------------------
// Householder QR with column pivoting of the small factor, columns never moved
// ---------------------------------------------------------------------------
struct QrcpArgs {
    double* At; int64_t ldm; int m; int nrows; int max_steps; double tol2;   // nrows: rows of the factor that can be non-zero
    double* cand_val; int* cand_idx;      // [2][G]
    int* done_step;                       // [ldm] step at which a column was eliminated (-1: not yet)
    double* diagval;                      // [m] beta of each step
    int* r_out;
};

__device__ __forceinline__ void block_sum8(double (&v)[8], double (*red)[8], double* out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double s = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (tid < 8) {
        double s = 0.;
#pragma unroll
        for (int w = 0; w < QT / 32; ++w) s += red[w][tid];
        out[tid] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ void block_argmax(double& v, int& i, double* rv, int* ri) {
    const int tid = threadIdx.x;
    rv[tid] = v; ri[tid] = i; __syncthreads();
    for (int s = QT / 2; s > 0; s >>= 1) {
        if (tid < s) {
            const double ov = rv[tid + s]; const int oi = ri[tid + s];
            if (ov > rv[tid] || (ov == rv[tid] && oi < ri[tid])) { rv[tid] = ov; ri[tid] = oi; }
        }
        __syncthreads();
    }
    v = rv[0]; i = ri[0];
    __syncthreads();
}

__global__ void __launch_bounds__(QT, 1) qrcp_kernel(QrcpArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double qsm[];
    const int m = a.m, nr = a.nrows; const int64_t ldm = a.ldm;
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const int ngroups = (int)(ldm / 8);
    const int mg = b < ngroups ? (ngroups - b + G - 1) / G : 0;      // my column groups: b, b + G, ...
    double* xs = qsm;                                   // [nr] the pivot column below the diagonal row
    double* cn2 = qsm + (nr + 7) / 8 * 8;                // [mg * 8] squared norms of my columns below the current row
    __shared__ int active[64 * 8];                      // my columns still to be eliminated (mg <= 64)
    __shared__ double red[QT / 32][8];
    __shared__ double out8[8];
    __shared__ double red16[QT / 32][QIB];
    __shared__ double out16[QIB], wsh16[QIB];
    __shared__ int act16[QIB];
    __shared__ double rv[QT]; __shared__ int ri[QT];
    double* __restrict__ At = a.At;

    // ---- prologue: norms of my columns over all rows
    for (int j = 0; j < mg; ++j) {
        const int k0 = (b + j * G) * 8;
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.;
        for (int r = tid; r < nr; r += QT) {
            const double2* p = reinterpret_cast<const double2*>(At + (int64_t)r * ldm + k0);
#pragma unroll
            for (int q = 0; q < 4; ++q) { double2 v = p[q]; acc[2 * q] += v.x * v.x; acc[2 * q + 1] += v.y * v.y; }
        }
        block_sum8(acc, red, out8);
        if (tid < 8) {
            const bool real = k0 + tid < m;
            cn2[j * 8 + tid] = real ? out8[tid] : -1.;
            active[j * 8 + tid] = real ? 1 : 0;
            a.done_step[k0 + tid] = -1;
        }
        __syncthreads();
    }
    __syncthreads();

    double first = 0.;
    int par = 0, c = 0;
    for (; c < a.max_steps; ++c) {
        // ---- pivot: largest remaining column norm (ties -> lowest column index)
        double bv = -1.; int bi = 0x7fffffff;
        for (int e = tid; e < mg * 8; e += QT) {
            if (active[e]) {
                const double v = cn2[e]; const int k = (b + (e / 8) * G) * 8 + (e % 8);
                if (v > bv || (v == bv && k < bi)) { bv = v; bi = k; }
            }
        }
        block_argmax(bv, bi, rv, ri);
        if (tid == 0) { a.cand_val[par * G + b] = bv; a.cand_idx[par * G + b] = bi; }
        grid.sync();
        bv = -1.; bi = 0x7fffffff;
        for (int e = tid; e < G; e += QT) {
            const double v = __ldcg(a.cand_val + par * G + e); const int k = __ldcg(a.cand_idx + par * G + e);
            if (v > bv || (v == bv && k < bi)) { bv = v; bi = k; }
        }
        block_argmax(bv, bi, rv, ri);
        par ^= 1;
        const int p = bi; const double pv = bv;
        if (c == 0) first = pv;
        if (!(pv > a.tol2 * first) || !(pv > 0.)) break;            // uniform over the grid

        // ---- reflector from column p, rows c..m-1
        const int len = nr - c;
        // column p was written by its owner CTA before the grid barrier above (grid.sync orders those writes before these
        // reads).  Plain loads: forcing them through L2 (ld.cg) made the kernel 3x slower at m = 4096 (214 vs 75 ms)
        for (int r = tid; r < len; r += QT) xs[r] = At[(int64_t)(c + r) * ldm + p];
        __syncthreads();
        double part = 0.;
        for (int r = 1 + tid; r < len; r += QT) part += xs[r] * xs[r];
        rv[tid] = part; __syncthreads();
        for (int s = QT / 2; s > 0; s >>= 1) { if (tid < s) rv[tid] += rv[tid + s]; __syncthreads(); }
        const double xn2 = rv[0], alpha = xs[0];
        __syncthreads();
        double beta = alpha, tau = 0., scale = 0.;
        if (xn2 > 0.) {
            beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
            tau = (beta - alpha) / beta;
            scale = 1. / (alpha - beta);
        }
        const bool mine = (p / 8) % G == b;
        if (mine && tid == 0) {
            a.diagval[c] = beta; a.done_step[p] = c;
            const int e = ((p / 8 - b) / G) * 8 + (p % 8);
            active[e] = 0; cn2[e] = -1.;
        }
        __syncthreads();

        // ---- update my remaining columns, recompute their norms below row c.  Two column groups (16 columns) per pass:
        // half the passes over the rows and half the block reductions of a group-at-a-time loop (m = 4096: 4 groups per CTA)
        for (int j0 = 0; j0 < mg; j0 += 2) {
            const bool two = j0 + 1 < mg;
            const int ka = (b + j0 * G) * 8, kb = two ? (b + (j0 + 1) * G) * 8 : ka;
            int any = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) any |= active[j0 * 8 + q] | (two ? active[(j0 + 1) * 8 + q] : 0);
            if (!any) continue;                                        // uniform (shared memory)
            double acc[QIB];
#pragma unroll
            for (int q = 0; q < QIB; ++q) acc[q] = 0.;
            for (int r = c + 1 + tid; r < nr; r += QT) {
                const double xr = xs[r - c];
                const double2* pa = reinterpret_cast<const double2*>(At + (int64_t)r * ldm + ka);
                const double2* pb = reinterpret_cast<const double2*>(At + (int64_t)r * ldm + kb);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 va = pa[q], vb = pb[q];
                    acc[2 * q] += xr * va.x; acc[2 * q + 1] += xr * va.y;
                    acc[8 + 2 * q] += xr * vb.x; acc[8 + 2 * q + 1] += xr * vb.y;
                }
            }
            block_sum16(acc, red16, out16);
            if (tid < QIB) {
                const int e = (tid < 8 ? j0 : j0 + 1) * 8 + (tid & 7);
                const int col = (tid < 8 ? ka : kb) + (tid & 7);
                const bool act = (tid < 8 || two) && active[e];
                const double rc = At[(int64_t)c * ldm + col];
                const double wv = rc + out16[tid] * scale;
                wsh16[tid] = wv; act16[tid] = act ? 1 : 0;
                if (act) At[(int64_t)c * ldm + col] = rc - tau * wv;
            }
            __syncthreads();
            double w[QIB]; int am = 0;
#pragma unroll
            for (int q = 0; q < QIB; ++q) { w[q] = wsh16[q]; am |= act16[q] << q; }
#pragma unroll
            for (int q = 0; q < QIB; ++q) acc[q] = 0.;
            if (am == 0xffff) {                                        // all 16 columns active: 16-byte accesses
                for (int r = c + 1 + tid; r < nr; r += QT) {
                    const double tv = tau * (xs[r - c] * scale);
                    double2* pa = reinterpret_cast<double2*>(At + (int64_t)r * ldm + ka);
                    double2* pb = reinterpret_cast<double2*>(At + (int64_t)r * ldm + kb);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        double2 va = pa[q], vb = pb[q];
                        va.x -= tv * w[2 * q]; va.y -= tv * w[2 * q + 1];
                        vb.x -= tv * w[8 + 2 * q]; vb.y -= tv * w[8 + 2 * q + 1];
                        pa[q] = va; pb[q] = vb;
                        acc[2 * q] += va.x * va.x; acc[2 * q + 1] += va.y * va.y;
                        acc[8 + 2 * q] += vb.x * vb.x; acc[8 + 2 * q + 1] += vb.y * vb.y;
                    }
                }
            } else {
                for (int r = c + 1 + tid; r < nr; r += QT) {
                    const double tv = tau * (xs[r - c] * scale);
                    double* pa = At + (int64_t)r * ldm + ka;
                    double* pb = At + (int64_t)r * ldm + kb;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (am & (1 << q)) { const double v = pa[q] - tv * w[q]; pa[q] = v; acc[q] += v * v; }
                        if (am & (1 << (8 + q))) { const double v = pb[q] - tv * w[8 + q]; pb[q] = v; acc[8 + q] += v * v; }
                    }
                }
            }
            block_sum16(acc, red16, out16);
            if (tid < QIB && act16[tid]) cn2[(tid < 8 ? j0 : j0 + 1) * 8 + (tid & 7)] = out16[tid];
            __syncthreads();
        }
    }
    if (b == 0 && tid == 0) *a.r_out = c;
}

// put the betas on the diagonal positions, zero what lies below them and every row >= r
__global__ void qrcp_cleanup_kernel(double* __restrict__ At, int64_t ldm, int64_t rows_total, int m,
                                    const int* __restrict__ done_step, const double* __restrict__ diagval,
                                    const int* __restrict__ r_ptr) {
    const int r_fin = *r_ptr;
    const int64_t total = rows_total * ldm;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / ldm; const int k = (int)(idx % ldm);
        if (r >= r_fin || k >= m) { At[idx] = 0.; continue; }
        const int s = done_step[k];
        if (s >= 0 && s < r_fin) {
            if (r == s) At[idx] = diagval[s];
            else if (r > s) At[idx] = 0.;
        }
    }
}

int coop_grid(pb_ctx* ctx, const void* kernel, size_t smem, int want, int* grid_out) {
    int occ = 0;
    PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, QT, smem));
    if (occ < 1) PB_FAIL(ctx, PB_ERR_CUDA, "tsqr: cooperative kernel does not fit an SM (%zu bytes of shared memory)", smem);
    *grid_out = std::max(1, std::min(want, ctx->sm_count));
    return PB_OK;
}

}  // namespace

int pbi_gemm_nn_sub(pb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t m, const double* B, int64_t ldb,
                    int64_t k, double* Y, int64_t ldy);

namespace {
constexpr int64_t RS_MAX = 1400;                 // rows of an inner panel staged per CTA (1400 x 18 doubles = 197 KB)
int qr_leaf(pb_ctx* ctx, const double* A, int64_t rows, int64_t m, int64_t lda, double* R, int64_t ldr, int64_t rows_out,
            bool rr, int64_t* f_rows);
}

// Factor F of A (rows x m, lda): F^T F = A^T A to working accuracy, written to R (ld ldr, rows_out >= m rows
// available, all of them written -- zero below the factor).
//   rr == false: F = the m x m upper-triangular R of the blocked Householder QR (LAPACK dgeqrf's R up to row signs).
//   rr == true:  rank-revealing (block pivoting + early termination, see above): *f_rows <= ceil64(m) rows, dense, in the
//                input's column order.
// Slabs whose working copy would exceed 16 GiB are cut into equal row blocks of at most sm_count x 1400 rows -- TSQR
// leaves inside the GPU: F_i = qr(A_i), then the QR of the stacked factors.  The working copy is then one leaf, not the
// whole slab (6.8 GB instead of 41 GB at C5), every leaf runs the shared-memory panel path, and the stack adds
// n_leaves m / rows of the flops (2 % at 1.25 M x 4096).  Measured time-neutral at 1.25 M x 4096 (2.20 vs 2.24 s) and
// 4 % slower at 250 000 x 4096 (two leaves), hence the memory criterion.
int pbi_qr_factor(pb_ctx* ctx, const double* A, int64_t rows, int64_t m, int64_t lda, double* R, int64_t ldr, int64_t rows_out,
                  bool rr, int64_t* f_rows) {
    if (rows <= 0 || m <= 0 || lda < m || ldr < m || rows_out < m) PB_FAIL(ctx, PB_ERR_SHAPE, "tsqr: bad shape");
    if (m > 8192) PB_FAIL(ctx, PB_ERR_SHAPE, "tsqr: %lld columns exceed 8192", (long long)m);
    const char* le = getenv("PB_QR_LEAVES");            // 0: never, 1 (default): by the memory criterion, 2: whenever possible
    const int leaves_env = le ? atoi(le) : 1;
    const int64_t leaf_max = (int64_t)ctx->sm_count * RS_MAX;
    const int64_t n_leaves = (rows + leaf_max - 1) / leaf_max;
    const bool big = (double)rows * (double)((m + QIB - 1) / QIB * QIB) * 8. > 16. * 1073741824.;
    int64_t fr = m;
    if (!(leaves_env == 2 || (leaves_env == 1 && big)) || n_leaves < 2 || rows < 4 * m || n_leaves * m > leaf_max) {
        PB_TRY(qr_leaf(ctx, A, rows, m, lda, R, ldr, rows_out, rr, &fr));
        if (f_rows) *f_rows = fr;
        return PB_OK;
    }
    // every leaf zero-fills m rows from its offset, so the stack needs one factor of slack
    PB_TRY(pb_reserve(ctx, &ctx->qr_stack, &ctx->qr_stack_cap, (size_t)(n_leaves + 1) * m * m));
    const int64_t per = (rows + n_leaves - 1) / n_leaves;
    int64_t off = 0;
    for (int64_t i = 0; i < n_leaves; ++i) {
        const int64_t r0 = i * per, nr = std::min(per, rows - r0);
        PB_TRY(qr_leaf(ctx, A + r0 * lda, nr, m, lda, ctx->qr_stack + (size_t)off * m, m, m, rr, &fr));
        off += std::min(fr, m);
    }
    PB_TRY(qr_leaf(ctx, ctx->qr_stack, off, m, m, R, ldr, rows_out, rr, &fr));
    if (f_rows) *f_rows = fr;
    return PB_OK;
}

int pbi_qr_r(pb_ctx* ctx, const double* A, int64_t rows, int64_t m, int64_t lda, double* R, int64_t ldr, int64_t rows_out) {
    return pbi_qr_factor(ctx, A, rows, m, lda, R, ldr, rows_out, false, nullptr);
}

namespace {
// exact residual norms^2 of the columns at positions >= J over the rows >= J
int qr_exact_norms(pb_ctx* ctx, const double* W, int64_t ldw, int64_t rowsW, int J, int mp, double* part, double* cn2) {
    const int ncols = mp - J;
    const int ctile = (ncols + CN_COLS - 1) / CN_COLS;
    int nrb = std::max(1, std::min<int>((ctx->sm_count * 8 + ctile - 1) / ctile, (int)((rowsW - J + 63) / 64)));
    qr_colnorm_part_kernel<<<dim3(ctile, nrb), 128, 0, ctx->stream>>>(W, ldw, J, rowsW, J, ncols, part, mp);
    PB_LAUNCH_CHECK(ctx);
    qr_colnorm_reduce_kernel<<<(ncols + 255) / 256, 256, 0, ctx->stream>>>(part, nrb, mp, ncols, cn2 + J);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

int qr_leaf(pb_ctx* ctx, const double* A, int64_t rows, int64_t m, int64_t lda, double* R, int64_t ldr, int64_t rows_out,
            bool rr, int64_t* f_rows) {
    const int64_t mp = (m + QIB - 1) / QIB * QIB;            // padded column count (zero columns reflect to nothing)
    const int64_t ldw = mp;
    const int64_t rowsW = std::max(rows, mp);
    PB_TRY(pb_reserve(ctx, &ctx->qr_W, &ctx->qr_W_cap, (size_t)rowsW * ldw));
    const int Gmax = ctx->sm_count;
    // scratch: tau [mp] | part [2*G*16] | diag [32] | zpart [G*ZE] | zfin [ZE] | [S | Z] [64*mp] | T^T [64*64]
    //          | rank-revealing mode: cn2 [mp] | first [8] | norm partials [8 sm_count x mp] | colperm (ints) [mp]
    const int nrb_max = Gmax * 8;
    const size_t n_scr = (size_t)mp + 2 * (size_t)Gmax * QIB + 2 * QIB + (size_t)Gmax * ZE + ZE + (size_t)QNB * mp + QNB * QNB
                         + (rr ? (size_t)mp + 8 + (size_t)nrb_max * mp + (size_t)(mp + 1) / 2 + 8 : 0);
    PB_TRY(pb_reserve(ctx, &ctx->qr_scr, &ctx->qr_scr_cap, n_scr));
    double* tau = ctx->qr_scr; double* part = tau + mp; double* diag = part + 2 * (size_t)Gmax * QIB;
    double* zpart = diag + 2 * QIB; double* zfin = zpart + (size_t)Gmax * ZE; double* Zb = zfin + ZE;
    double* Tt = Zb + (size_t)QNB * mp;
    double* cn2 = Tt + QNB * QNB; double* first = cn2 + mp; double* npart = first + 8;
    int* colperm = rr ? reinterpret_cast<int*>(npart + (size_t)nrb_max * mp) : nullptr;
    int* piv_out = ctx->d_int + 64; int* piv_host = ctx->h_int + 64;        // [0] stop [1] want exact [2] swaps [4..] pairs
    double* W = ctx->qr_W;
    PB_CUDA(ctx, cudaMemset2DAsync(R, ldr * sizeof(double), 0, m * sizeof(double), rows_out, ctx->stream));
    const int cp_grid = ctx->sm_count * 8;
    static const int stage_env = getenv("PB_QR_STAGE") ? atoi(getenv("PB_QR_STAGE")) : 1;
    static const int pivot_env = getenv("PB_QR_PIVOT") ? atoi(getenv("PB_QR_PIVOT")) : 1;
    static const int fold_env = getenv("PB_QR_FOLD") ? atoi(getenv("PB_QR_FOLD")) : 1;
    const size_t tt_smem = sizeof(double) * 2 * QNB * (QNB + 1);
    PB_TRY(pbi_smem_optin(ctx, (const void*)qr_tt_kernel, (size_t)(tt_smem)));
    bool folded = false;                                              // panel 0's pivoting was applied by the copy
    if (rr) {
        qr_iota_kernel<<<(unsigned)((mp + 255) / 256), 256, 0, ctx->stream>>>(colperm, (int)mp);
        PB_LAUNCH_CHECK(ctx);
        PB_TRY(pbi_smem_optin(ctx, (const void*)qr_pivot_kernel, (size_t)((sizeof(double) * mp))));
        // First panel: the column norms come from A itself (the same values the copy would hold), the pivot kernel
        // leaves position -> source column in colperm, and the working copy is made through that map: one pass over
        // the slab less than copy + swap (C2: 0.32 ms of a 7.5 ms step)
        if (pivot_env && fold_env && mp > QNB && (m % 2 == 0) && (lda % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0)) {
            PB_CUDA(ctx, cudaMemsetAsync(cn2, 0, sizeof(double) * mp, ctx->stream));
            const int ctile = (int)((m + CN_COLS - 1) / CN_COLS);
            const int nrb = std::max(1, std::min<int>((ctx->sm_count * 8 + ctile - 1) / ctile, (int)((rows + 63) / 64)));
            qr_colnorm_part_kernel<<<dim3(ctile, nrb), 128, 0, ctx->stream>>>(A, lda, 0, rows, 0, (int)m, npart, mp);
            PB_LAUNCH_CHECK(ctx);
            qr_colnorm_reduce_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(npart, nrb, mp, (int)m, cn2);
            PB_LAUNCH_CHECK(ctx);
            QrPivArgs pa{cn2, colperm, 0, (int)std::min<int64_t>(QNB, mp), (int)mp, 0., first, 1, pivot_env, piv_out};
            qr_pivot_kernel<<<1, PIVT, sizeof(double) * mp, ctx->stream>>>(pa);
            PB_LAUNCH_CHECK(ctx);
            PB_CUDA(ctx, cudaMemcpyAsync(piv_host, piv_out, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (!piv_host[0] && !piv_host[1]) {                       // (a zero matrix stops here: the plain path reports it)
                qr_copy_pad_perm_kernel<<<cp_grid, 256, 0, ctx->stream>>>(A, lda, rows, m, W, ldw, rowsW, colperm);
                PB_LAUNCH_CHECK(ctx);
                folded = true;
            }
        }
    }
    if (!folded) {
        qr_copy_pad_kernel<<<cp_grid, 256, 0, ctx->stream>>>(A, lda, rows, m, W, ldw, rowsW);
        PB_LAUNCH_CHECK(ctx);
    }
    int64_t K = mp;                                                   // rows of the factor (panels actually run)
    int panels = 0;
    for (int64_t J = 0; J < mp; J += QNB) {
        const int nbo = (int)std::min<int64_t>(QNB, mp - J);
        if (rr && !(folded && J == 0)) {
            // residual must fall below the rounding noise the panels run so far have put into the kept rows
            const double tol = 16. * DBL_EPSILON * sqrt((double)std::max(1, panels));
            int exact = 0;
            if (J == 0) { PB_TRY(qr_exact_norms(ctx, W, ldw, rowsW, 0, (int)mp, npart, cn2)); exact = 1; }
            for (;;) {
                QrPivArgs pa{cn2, colperm, (int)J, nbo, (int)mp, tol * tol, first, exact, pivot_env, piv_out};
                qr_pivot_kernel<<<1, PIVT, sizeof(double) * (mp - J), ctx->stream>>>(pa);
                PB_LAUNCH_CHECK(ctx);
                PB_CUDA(ctx, cudaMemcpyAsync(piv_host, piv_out, sizeof(int) * (4 + 2 * QNB), cudaMemcpyDeviceToHost, ctx->stream));
                PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                if (piv_host[1] && !exact) { PB_TRY(qr_exact_norms(ctx, W, ldw, rowsW, (int)J, (int)mp, npart, cn2)); exact = 1; continue; }
                break;
            }
            static const int dbg = getenv("PB_QR_DEBUG") ? atoi(getenv("PB_QR_DEBUG")) : 0;
            if (dbg) {
                double hv[2] = {0., 0.};
                std::vector<double> hc((size_t)(mp - J));
                cudaMemcpy(hv, first, sizeof(double), cudaMemcpyDeviceToHost);
                cudaMemcpy(hc.data(), cn2 + J, sizeof(double) * (mp - J), cudaMemcpyDeviceToHost);
                double mxv = 0.; for (double v : hc) mxv = std::max(mxv, v);
                fprintf(stderr, "[qr] rows %lld J %lld exact %d max/first %.3e (sqrt %.3e) tol %.3e stop %d swaps %d\n", (long long)rows,
                        (long long)J, exact, mxv / hv[0], sqrt(mxv / hv[0]), tol, piv_host[0], piv_host[2]);
            }
            if (piv_host[0]) { K = J; break; }
            if (piv_host[2] > 0) {
                const int ns = piv_host[2];
                qr_swap_cols_kernel<<<(unsigned)std::min<int64_t>(cp_grid * 4, (rowsW * ns + 255) / 256), 256, 0, ctx->stream>>>(
                    W, ldw, rowsW, piv_out + 4, ns);
                PB_LAUNCH_CHECK(ctx);
            }
        }
        int grid = std::max(1, std::min<int>((int)((rowsW - J + 127) / 128), ctx->sm_count));
        int64_t rpc = (rowsW - J + grid - 1) / grid; rpc = (rpc + 7) / 8 * 8;       // as the kernel computes it
        const int use_smem = (stage_env && rpc <= RS_MAX) ? 1 : 0;
        const size_t smem = use_smem ? sizeof(double) * (size_t)rpc * QP : 0;
        PB_TRY(pbi_smem_optin(ctx, (const void*)qr_panel_kernel, (size_t)(smem)));
        PB_TRY(coop_grid(ctx, (const void*)qr_panel_kernel, smem, grid, &grid));
        QrPanelArgs pa{W, ldw, rowsW, (int)J, nbo, tau + J, part, diag, zpart, zfin, use_smem};
        void* args[] = {&pa};
        PB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)qr_panel_kernel, dim3(grid), dim3(QT), args, smem, ctx->stream));
        ctx->launches++;
        ++panels;
        {   // Householder flop count of this panel (factorisation + trailing update), over the rows still active
            const double ra = (double)std::max<int64_t>(0, rows - J), nt = (double)(mp - J - nbo);
            ctx->qr_flops += 2. * ra * nbo * nbo + 4. * ra * nbo * nt;
        }
        qr_save_r_make_y_kernel<<<(unsigned)std::min<int64_t>(cp_grid, ((J + nbo) * nbo + 255) / 256), 256, 0, ctx->stream>>>(
            W, ldw, (int)J, nbo, m, R, ldr, colperm);
        PB_LAUNCH_CHECK(ctx);
        const int64_t ntrail = mp - J - nbo;
        if (ntrail <= 0) break;
        const double* Y = W + J;                                      // the panel columns are the explicit block Y now
        const int64_t ldz = nbo + ntrail;
        PB_TRY(pbi_gemm_tn(ctx, Y, ldw, nbo, Y, ldw, ldz, rowsW, Zb, ldz, 0));                    // [Y^T Y | Y^T C]
        qr_tt_kernel<<<1, 4 * QNB, tt_smem, ctx->stream>>>(Zb, ldz, nbo, tau + J, Tt);
        PB_LAUNCH_CHECK(ctx);
        qr_applyt_kernel<<<(unsigned)((ntrail + QAT - 1) / QAT), QT, 0, ctx->stream>>>(Zb + nbo, ldz, ntrail, Tt, nbo);   // -X = -T^T Z
        PB_LAUNCH_CHECK(ctx);
        PB_TRY(pbi_gemm_nn_sub(ctx, Y, ldw, rowsW, nbo, Zb + nbo, ldz, ntrail, W + J + nbo, ldw)); // C += Y (-X)
        if (rr) {
            qr_downdate_kernel<<<(unsigned)((ntrail + 127) / 128), 128, 0, ctx->stream>>>(W, ldw, (int)J, nbo, (int)mp, cn2);
            PB_LAUNCH_CHECK(ctx);
        }
    }
    if (rr && K < mp && K > 0) {      // stopped early: the rows of R12 that belong to the columns never factored
        qr_gather_tail_kernel<<<(unsigned)std::min<int64_t>(cp_grid, (K * (mp - K) + 255) / 256), 256, 0, ctx->stream>>>(
            W, ldw, (int)K, (int)mp, m, colperm, R, ldr);
        PB_LAUNCH_CHECK(ctx);
    }
    ctx->qr_panels += panels;
    if (f_rows) *f_rows = std::min<int64_t>(K, m);
    return PB_OK;
}
}  // namespace

// At (rows_total x ldm, zero padded) = upper triangle of R (m x m, ld ldr)
int pbi_qr_load_factor(pb_ctx* ctx, const double* R, int64_t ldr, int64_t m, double* At, int64_t ldm, int64_t rows_total) {
    qr_extract_r_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(R, ldr, m, At, ldm, rows_total);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}
// At (rows_total x ldm, zero padded) = the dense factor F (f_rows x m, ld ldf)
int pbi_qr_load_dense(pb_ctx* ctx, const double* F, int64_t ldf, int64_t f_rows, int64_t m, double* At, int64_t ldm, int64_t rows_total) {
    qr_copy_pad_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(F, ldf, f_rows, m, At, ldm, rows_total);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

// Rank-revealing QR with column pivoting of the factor held in At (m rows x ldm, rows = vectors after the call; only the
// first nrows rows can be non-zero -- nrows = m for a triangular R, fewer for a truncated factor).
// tol_rel: stop when the largest remaining column norm <= tol_rel * the first pivot norm (0: all m steps).
// The number of vectors r is left in ctx->d_int[0] / returned.
int pbi_qrcp_factor(pb_ctx* ctx, double* At, int64_t m, int64_t nrows, int64_t ldm, int64_t rows_total, double tol_rel, int* r_out) {
    if (nrows <= 0 || nrows > m) nrows = m;
    const int Gmax = ctx->sm_count;
    const int ngroups = (int)(ldm / 8);
    int grid = std::min(Gmax, ngroups);
    const int mg_max = (ngroups + grid - 1) / grid;
    if (mg_max > 64) PB_FAIL(ctx, PB_ERR_SHAPE, "tsqr: %lld columns exceed the pivoted QR's limit", (long long)m);
    const size_t smem = sizeof(double) * ((size_t)(nrows + 7) / 8 * 8 + (size_t)mg_max * 8);
    PB_TRY(pbi_smem_optin(ctx, (const void*)qrcp_kernel, (size_t)(smem)));
    PB_TRY(coop_grid(ctx, (const void*)qrcp_kernel, smem, grid, &grid));
    // scratch: cand_val [2*G] | diagval [m] | then ints: cand_idx [2*G] | done_step [ldm]
    const size_t n_dbl = 2 * (size_t)Gmax + (size_t)m + 8;
    const size_t n_int = 2 * (size_t)Gmax + (size_t)ldm + 8;
    PB_TRY(pb_reserve(ctx, &ctx->qr_scr2, &ctx->qr_scr2_cap, n_dbl + (n_int + 1) / 2));
    double* cand_val = ctx->qr_scr2; double* diagval = cand_val + 2 * (size_t)Gmax;
    int* cand_idx = reinterpret_cast<int*>(diagval + m + 8); int* done_step = cand_idx + 2 * (size_t)Gmax;
    QrcpArgs qa{At, ldm, (int)m, (int)nrows, (int)nrows, tol_rel * tol_rel, cand_val, cand_idx, done_step, diagval, ctx->d_int + 0};
    void* args[] = {&qa};
    PB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)qrcp_kernel, dim3(grid), dim3(QT), args, smem, ctx->stream));
    ctx->launches++;
    qrcp_cleanup_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(At, ldm, rows_total, (int)m, done_step, diagval, ctx->d_int + 0);
    PB_LAUNCH_CHECK(ctx);
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->h_int, ctx->d_int, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *r_out = ctx->h_int[0];
    return PB_OK;
}
