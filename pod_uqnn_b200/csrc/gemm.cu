// fp64 tall-skinny products on the DMMA pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4).
//
//   TN:  C (ka x kb) = A^T B, A (rows x ka), B (rows x kb) row-major, rows >> ka,kb.
//        The Gram U^T U (sym = 1, only tiles j >= i), the projection V^T U, Y^T Y.
//        Reduction index = the long, strided one -> split-K over row ranges with a
//        fixed-order second-stage sum (deterministic, run-to-run identical ranks).
//   NN:  Y (rows x k) = A (rows x m) B (m x k).  Basis V = U (W S^-1), rotation Y = U W,
//        reconstruction V v^T (+ fused pod_sig epilogue).
//
// Staging: cp.async (LDGSTS) 16-byte copies into a 4-stage shared-memory ring.
// Shared rows are padded by 4 doubles so that the 8x4 / 4x8 DMMA fragment reads
// (4 k-rows x 4 consecutive columns per half-warp) hit 16 distinct 8-byte banks.
// tcgen05 has no fp64 kind; DMMA through mma.sync is the fp64 tensor path on sm_100a.
#include "common.cuh"

#include "tile_ops.cuh"

namespace {
using namespace pbt;

// ---------------------------------------------------------------------------
// TN kernel
// ---------------------------------------------------------------------------
template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, bool ALIGN16>
__global__ void __launch_bounds__(NTHREADS, (BM <= 32) ? 2 : 1)     // skinny (HBM-bound) tiles: two CTAs per SM
gemm_tn_kernel(const double* __restrict__ A, int64_t lda, int ka,
               const double* __restrict__ B, int64_t ldb, int kb,
               int64_t rows, int64_t rows_per_split,
               double* __restrict__ out, int64_t ldo, int64_t split_stride,
               int tiles_m, int tiles_n, int sym, int64_t bsA, int64_t bsB, int64_t bsOut) {
    static_assert(WARPS_M * WARPS_N == NTHREADS / 32, "8 warps");
    A += (int64_t)blockIdx.z * bsA; B += (int64_t)blockIdx.z * bsB; out += (int64_t)blockIdx.z * bsOut;   // batch
    constexpr int PA = BM + 4, PB = BN + 4;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int TM = WM / 8, TNN = WN / 8;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                              // [STAGES][BK][PA]
    double* Bs = smem + STAGES * BK * PA;           // [STAGES][BK][PB]

    int tile = blockIdx.x, ti, tj;
    if (sym) {           // linear id -> (ti, tj >= ti)
        ti = 0; int rem = tile;
        while (rem >= tiles_n - ti) { rem -= tiles_n - ti; ++ti; }
        tj = ti + rem;
    } else { ti = tile / tiles_n; tj = tile % tiles_n; }
    (void)tiles_m;
    const int64_t i0 = (int64_t)ti * BM, j0 = (int64_t)tj * BN;
    const int split = blockIdx.y;
    const int64_t r_begin = (int64_t)split * rows_per_split;
    const int64_t r_end = min(rows, r_begin + rows_per_split);
    const int nk = (int)((r_end - r_begin + BK - 1) / BK);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int lk = lane & 3, lg = lane >> 2;

    double acc[TM][TNN][2];
#pragma unroll
    for (int t = 0; t < TM; ++t)
#pragma unroll
        for (int u = 0; u < TNN; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            int64_t r0 = r_begin + (int64_t)s * BK;
            load_tile<BK, BM, PA, ALIGN16>(As + s * BK * PA, A, lda, r0, r_end, i0, ka, tid);
            load_tile<BK, BN, PB, ALIGN16>(Bs + s * BK * PB, B, ldb, r0, r_end, j0, kb, tid);
        }
        cp_async_commit();
    }

    for (int k = 0; k < nk; ++k) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // prefetch stage k + STAGES - 1 into the slot freed at iteration k - 1
            int kn = k + STAGES - 1;
            if (kn < nk) {
                int s = kn % STAGES;
                int64_t r0 = r_begin + (int64_t)kn * BK;
                load_tile<BK, BM, PA, ALIGN16>(As + s * BK * PA, A, lda, r0, r_end, i0, ka, tid);
                load_tile<BK, BN, PB, ALIGN16>(Bs + s * BK * PB, B, ldb, r0, r_end, j0, kb, tid);
            }
            cp_async_commit();
        }
        const double* as = As + (k % STAGES) * BK * PA + wm * WM + lg;
        const double* bs = Bs + (k % STAGES) * BK * PB + wn * WN + lg;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[TM], b[TNN];
#pragma unroll
            for (int t = 0; t < TM; ++t) a[t] = as[(kk * 4 + lk) * PA + t * 8];
#pragma unroll
            for (int u = 0; u < TNN; ++u) b[u] = bs[(kk * 4 + lk) * PB + u * 8];
#pragma unroll
            for (int t = 0; t < TM; ++t)
#pragma unroll
                for (int u = 0; u < TNN; ++u) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
        }
    }
    cp_async_wait<0>();

    double* o = out + (int64_t)split * split_stride;
#pragma unroll
    for (int t = 0; t < TM; ++t) {
        int64_t i = i0 + wm * WM + t * 8 + lg;
        if (i >= ka) continue;
#pragma unroll
        for (int u = 0; u < TNN; ++u) {
            int64_t j = j0 + wn * WN + u * 8 + 2 * lk;
            if (j < kb)     o[i * ldo + j]     = acc[t][u][0];
            if (j + 1 < kb) o[i * ldo + j + 1] = acc[t][u][1];
        }
    }
}

// second stage: fixed-order sum over splits (+ mirror of the upper triangle when sym)
__global__ void reduce_splits_kernel(const double* __restrict__ ws, int splits, int64_t split_stride,
                                     int ka, int kb, int64_t ldw, double* __restrict__ C, int64_t ldc,
                                     int sym, int bm, int bn, int64_t bsW, int64_t bsC) {
    ws += (int64_t)blockIdx.y * bsW; C += (int64_t)blockIdx.y * bsC;      // batch
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)ka * kb) return;
    int i = (int)(idx / kb), j = (int)(idx % kb);
    int si = i, sj = j;
    // in sym mode tile (ti, tj) exists only for tj >= ti: elements of lower tiles come mirrored
    if (sym && (j / bn) < (i / bm)) { si = j; sj = i; }
    const double* p = ws + (int64_t)si * ldw + sj;
    double s = 0.;
    for (int k = 0; k < splits; ++k) s += p[(int64_t)k * split_stride];
    C[(int64_t)i * ldc + j] = s;
}

// make an (n x n) matrix exactly symmetric from its upper triangle (diagonal tiles of the
// sym product are computed in full; rounding is identical on both sides because the same
// products are accumulated in the same order, but we do not rely on it).
__global__ void symmetrize_kernel(double* __restrict__ C, int n, int64_t ldc, int64_t bsC) {
    C += (int64_t)blockIdx.y * bsC;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * n) return;
    int i = (int)(idx / n), j = (int)(idx % n);
    if (j < i) C[(int64_t)i * ldc + j] = C[(int64_t)j * ldc + i];
}

// ---------------------------------------------------------------------------
// NN kernel (BM = 128 rows per CTA)
// ---------------------------------------------------------------------------
template <int BN, int WARPS_M, int WARPS_N, int STAGES, bool ALIGN_A, bool ALIGN_B>
__global__ void __launch_bounds__(NTHREADS, (BN <= 32) ? 2 : 1)     // skinny (HBM-bound) tiles: two CTAs per SM
gemm_nn_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int m,
               const double* __restrict__ B, int64_t ldb, int k,
               double* __restrict__ Y, int64_t ldy,
               const double* __restrict__ Uref, int64_t ldu, double* __restrict__ sig_part,
               int64_t bsA, int64_t bsB, int64_t bsY, int subtract) {
    constexpr int BM = 128;
    A += (int64_t)blockIdx.z * bsA; B += (int64_t)blockIdx.z * bsB;
    if (Y) Y += (int64_t)blockIdx.z * bsY;                                   // batch
    static_assert(WARPS_M * WARPS_N == NTHREADS / 32, "8 warps");
    constexpr int PA = BK + 4, PB = BN + 4;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int TM = WM / 8, TNN = WN / 8;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                              // [STAGES][BM][PA]
    double* Bs = smem + STAGES * BM * PA;           // [STAGES][BK][PB]
    __shared__ double rowsum[WARPS_N][BM];

    const int64_t h0 = (int64_t)blockIdx.x * BM;
    const int64_t j0 = (int64_t)blockIdx.y * BN;
    const int nk = (m + BK - 1) / BK;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int lk = lane & 3, lg = lane >> 2;

    double acc[TM][TNN][2];
#pragma unroll
    for (int t = 0; t < TM; ++t)
#pragma unroll
        for (int u = 0; u < TNN; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            int64_t k0 = (int64_t)s * BK;
            load_tile<BM, BK, PA, ALIGN_A>(As + s * BM * PA, A, lda, h0, rows, k0, m, tid);
            load_tile<BK, BN, PB, ALIGN_B>(Bs + s * BK * PB, B, ldb, k0, m, j0, k, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int kn = kt + STAGES - 1;
            if (kn < nk) {
                int s = kn % STAGES;
                int64_t k0 = (int64_t)kn * BK;
                load_tile<BM, BK, PA, ALIGN_A>(As + s * BM * PA, A, lda, h0, rows, k0, m, tid);
                load_tile<BK, BN, PB, ALIGN_B>(Bs + s * BK * PB, B, ldb, k0, m, j0, k, tid);
            }
            cp_async_commit();
        }
        const double* as = As + (kt % STAGES) * BM * PA + (wm * WM + lg) * PA + lk;
        const double* bs = Bs + (kt % STAGES) * BK * PB + wn * WN + lg;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[TM], b[TNN];
#pragma unroll
            for (int t = 0; t < TM; ++t) a[t] = as[t * 8 * PA + kk * 4];
#pragma unroll
            for (int u = 0; u < TNN; ++u) b[u] = bs[(kk * 4 + lk) * PB + u * 8];
#pragma unroll
            for (int t = 0; t < TM; ++t)
#pragma unroll
                for (int u = 0; u < TNN; ++u) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
        }
    }
    cp_async_wait<0>();

    // epilogues: Uref with ldu > 0 -> pod_sig partial row sums; ldu < 0 -> Monte-Carlo moment accumulation
    // (Uref = running sum, sig_part = running sum of squares, both rows x k with leading dimension -ldu)
    const bool want_mom = (Uref != nullptr) && ldu < 0;
    const bool want_sig = (Uref != nullptr) && !want_mom;
    double* msum = const_cast<double*>(Uref); const int64_t ldm_ = -ldu;
#pragma unroll
    for (int t = 0; t < TM; ++t) {
        const int rl = wm * WM + t * 8 + lg;
        const int64_t h = h0 + rl;
        double rs = 0.;
        if (h < rows) {
#pragma unroll
            for (int u = 0; u < TNN; ++u) {
                int64_t j = j0 + wn * WN + u * 8 + 2 * lk;
                if (Y) {
                    if (subtract) {          // Y += A B (trailing update of the blocked Householder QR, qr.cu: B = -X)
                        if (j < k)     Y[h * ldy + j]     += acc[t][u][0];
                        if (j + 1 < k) Y[h * ldy + j + 1] += acc[t][u][1];
                    } else {
                        if (j < k)     Y[h * ldy + j]     = acc[t][u][0];
                        if (j + 1 < k) Y[h * ldy + j + 1] = acc[t][u][1];
                    }
                }
                if (want_mom) {
                    if (j < k)     { msum[h * ldm_ + j] += acc[t][u][0];     sig_part[h * ldm_ + j] += acc[t][u][0] * acc[t][u][0]; }
                    if (j + 1 < k) { msum[h * ldm_ + j + 1] += acc[t][u][1]; sig_part[h * ldm_ + j + 1] += acc[t][u][1] * acc[t][u][1]; }
                }
                if (want_sig) {
                    if (j < k)     rs += fabs(Uref[h * ldu + j] - acc[t][u][0]);
                    if (j + 1 < k) rs += fabs(Uref[h * ldu + j + 1] - acc[t][u][1]);
                }
            }
        }
        if (want_sig) {   // quad reduction (the 4 lanes of a row), fixed order
            rs += __shfl_xor_sync(0xffffffffu, rs, 1);
            rs += __shfl_xor_sync(0xffffffffu, rs, 2);
            if (lk == 0) rowsum[wn][rl] = rs;
        }
    }
    if (want_sig) {
        __syncthreads();
        if (tid < BM && h0 + tid < rows) {
            double s = 0.;
#pragma unroll
            for (int w = 0; w < WARPS_N; ++w) s += rowsum[w][tid];
            sig_part[(int64_t)blockIdx.y * rows + h0 + tid] = s;
        }
    }
}

// ---------------------------------------------------------------------------
// Reconstruction strip kernel: Y (rows x n) = A (rows x K) B (K x n) with K = n_L <= 64 and n wide -- the op is
// bound by the output / U traffic, not by flops.  One CTA owns a strip of 32 rows and sweeps the column tiles:
// the A strip is staged once, the B tile and the U tile of the NEXT column tile are in flight (cp.async) while the
// current tile is multiplied and compared, so HBM streams continuously; a CTA owns whole rows, so
// pod_sig[h] = mean_j |U - Y| / 2 is finished in the kernel (no partial buffer, no second pass).
// ---------------------------------------------------------------------------
template <int KP, int MS, bool ALIGN_A, bool ALIGN_B, bool ALIGN_U>
__global__ void __launch_bounds__(NTHREADS, (KP <= 32 && MS == 32) ? 2 : 1)
recon_strip_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int K,
                   const double* __restrict__ B, int64_t ldb, int64_t n,
                   double* __restrict__ Y, int64_t ldy,
                   const double* U, int64_t ldu, double* __restrict__ sig_out, int subtract) {
    // subtract != 0: Y = U + A B with U == Y allowed (the in-place trailing update C -= Y X of the blocked
    // Householder QR, qr.cu, which passes B = -X): the accumulators start from the staged copy of the U tile, so the
    // epilogue is stores only.
    // MS = rows per strip: 32 for the HBM-bound reconstruction; 64 for the QR update (K = 64 is compute bound and a
    // 32-row strip re-reads the B tiles from L2 twice as often as it can afford)
    constexpr int NT = 64;
    constexpr int RA = MS / 16;                // 8-row fragments per warp (2 x 4 warps, warp tile MS/2 x 16)
    constexpr int STG = (KP <= 16) ? 3 : 2;    // tiles in flight: 2 CTAs x 2 tiles x 16 KB keep HBM busy (n_L <= 16)
    constexpr int PA = KP + 4, PB = NT + 4;
    constexpr int PU = NT + 8;                 // the U tile is read back as double2 per (row, column pair): a row pitch of
                                               // 64 bytes mod 128 keeps the 8 lanes of a quarter-warp on distinct banks
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                         // [MS][PA]
    double* Bs = As + MS * PA;                 // [STG][KP][PB]
    double* Us = Bs + STG * KP * PB;           // [STG][MS][PU]
    __shared__ double rowsum[4][MS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int lk = lane & 3, lg = lane >> 2;
    const int64_t h0 = (int64_t)blockIdx.x * MS;
    const int n_tiles = (int)((n + NT - 1) / NT);
    const bool want_u = (U != nullptr);
    const bool want_sig = want_u && !subtract;
    const bool y_vec = Y && (ldy % 2 == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);

    load_tile<MS, KP, PA, ALIGN_A>(As, A, lda, h0, rows, 0, K, tid);
    auto prefetch = [&](int t) {
        const int64_t j0 = (int64_t)t * NT;
        load_tile<KP, NT, PB, ALIGN_B>(Bs + (t % STG) * KP * PB, B, ldb, 0, K, j0, n, tid);
        if (want_u) load_tile<MS, NT, PU, ALIGN_U>(Us + (t % STG) * MS * PU, U, ldu, h0, rows, j0, n, tid);
    };
#pragma unroll
    for (int t = 0; t < STG - 1; ++t) {
        if (t < n_tiles) prefetch(t);
        cp_async_commit();
    }
    double rs[RA];
#pragma unroll
    for (int a = 0; a < RA; ++a) rs[a] = 0.;
    for (int t = 0; t < n_tiles; ++t) {
        cp_async_wait<STG - 2>();
        __syncthreads();                       // tile t has landed; everyone is done with tile t - 1's buffers
        if (t + STG - 1 < n_tiles) prefetch(t + STG - 1);
        cp_async_commit();
        double acc[RA][2][2];
        const double* us = Us + (t % STG) * MS * PU;
        if (subtract) {      // in-place update: the accumulators start from the staged C tile, the caller passes B = -X
#pragma unroll
            for (int a = 0; a < RA; ++a)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const double2 uv = *reinterpret_cast<const double2*>(us + (wm * (MS / 2) + a * 8 + lg) * PU + wn * 16 + u * 8 + 2 * lk);
                    acc[a][u][0] = uv.x; acc[a][u][1] = uv.y;
                }
        } else {
#pragma unroll
            for (int a = 0; a < RA; ++a)
#pragma unroll
                for (int u = 0; u < 2; ++u) { acc[a][u][0] = 0.; acc[a][u][1] = 0.; }
        }
        const double* as = As + (wm * (MS / 2) + lg) * PA + lk;
        const double* bs = Bs + (t % STG) * KP * PB + wn * 16 + lg;
        if (K == KP) {       // full-depth product (the QR update): straight-line code, the fragment loads can run ahead
#pragma unroll
            for (int kk = 0; kk < KP / 4; ++kk) {
                double af[RA], bf[2];
#pragma unroll
                for (int a = 0; a < RA; ++a) af[a] = as[a * 8 * PA + kk * 4];
#pragma unroll
                for (int u = 0; u < 2; ++u) bf[u] = bs[(kk * 4 + lk) * PB + u * 8];
#pragma unroll
                for (int a = 0; a < RA; ++a)
#pragma unroll
                    for (int u = 0; u < 2; ++u) dmma884(acc[a][u][0], acc[a][u][1], af[a], bf[u]);
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < KP / 4; ++kk) {
                if (kk * 4 < K) {
                    double af[RA], bf[2];
#pragma unroll
                    for (int a = 0; a < RA; ++a) af[a] = as[a * 8 * PA + kk * 4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) bf[u] = bs[(kk * 4 + lk) * PB + u * 8];
#pragma unroll
                    for (int a = 0; a < RA; ++a)
#pragma unroll
                        for (int u = 0; u < 2; ++u) dmma884(acc[a][u][0], acc[a][u][1], af[a], bf[u]);
                }
            }
        }
        const int64_t j0 = (int64_t)t * NT;
#pragma unroll
        for (int a = 0; a < RA; ++a) {
            const int rl = wm * (MS / 2) + a * 8 + lg;
            const int64_t h = h0 + rl;
            if (h >= rows) continue;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int cl = wn * 16 + u * 8 + 2 * lk;
                const int64_t j = j0 + cl;
                if (Y) {
                    if (y_vec && j + 1 < n) *reinterpret_cast<double2*>(Y + h * ldy + j) = make_double2(acc[a][u][0], acc[a][u][1]);
                    else { if (j < n) Y[h * ldy + j] = acc[a][u][0]; if (j + 1 < n) Y[h * ldy + j + 1] = acc[a][u][1]; }
                }
                if (want_sig) {
                    const double2 uv = *reinterpret_cast<const double2*>(us + rl * PU + cl);
                    if (j < n)     rs[a] += fabs(uv.x - acc[a][u][0]);
                    if (j + 1 < n) rs[a] += fabs(uv.y - acc[a][u][1]);
                }
            }
        }
    }
    cp_async_wait<0>();
    if (want_sig) {
#pragma unroll
        for (int a = 0; a < RA; ++a) {
            double v = rs[a];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (lk == 0) rowsum[wn][wm * (MS / 2) + a * 8 + lg] = v;
        }
        __syncthreads();
        if (tid < MS && h0 + tid < rows)
            sig_out[h0 + tid] = (rowsum[0][tid] + rowsum[1][tid] + rowsum[2][tid] + rowsum[3][tid]) * (0.5 / (double)n);
    }
}

__global__ void sig_reduce_kernel(const double* __restrict__ part, int ntiles, int64_t rows,
                                  double inv2n, double* __restrict__ out) {
    int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= rows) return;
    double s = 0.;
    for (int t = 0; t < ntiles; ++t) s += part[(int64_t)t * rows + h];
    out[h] = s * inv2n;
}

__global__ void transpose_kernel(const double* __restrict__ in, int64_t rows, int64_t cols,
                                 int64_t ldin, double* __restrict__ out, int64_t ldout) {
    __shared__ double t[32][33];
    int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int64_t r = r0 + y, c = c0 + threadIdx.x;
        t[y][threadIdx.x] = (r < rows && c < cols) ? in[r * ldin + c] : 0.;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int64_t c = c0 + y, r = r0 + threadIdx.x;
        if (c < cols && r < rows) out[c * ldout + r] = t[threadIdx.x][y];
    }
}

template <typename K>
int set_smem(pb_ctx* ctx, K kernel, size_t bytes) {
    PB_TRY(pbi_smem_optin(ctx, (const void*)kernel, (size_t)(bytes)));
    return PB_OK;
}

inline bool aligned16(const void* p, int64_t ld) { return ((uintptr_t)p % 16 == 0) && (ld % 2 == 0); }

// choose the split count so that tiles*splits fills whole waves of the SMs
int choose_splits(int tiles, int64_t rows, int sms) {
    int64_t max_by_rows = rows / (BK * 32);      // >= 512 rows per split
    if (max_by_rows < 1) max_by_rows = 1;
    int best = 1; double best_score = -1.;
    for (int s = 1; s <= 96 && s <= max_by_rows; ++s) {
        int64_t ctas = (int64_t)tiles * s;
        int64_t waves = (ctas + sms - 1) / sms;
        double eff = (double)ctas / (double)(waves * sms);
        // prefer >= 2 waves of work per SM is not needed (equal CTAs); penalise many splits lightly
        double score = eff - 0.0015 * s;
        if (score > best_score) { best_score = score; best = s; }
    }
    return best;
}

}  // namespace
int pbi_gram_tma(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka, int64_t rows, int splits,
                 int64_t rows_per_split, double* ws, int64_t split_stride, bool time_it);
int pbi_gram_streamk(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka, int64_t rows, double* C, int64_t ldc,
                     bool time_it);
int pbi_qr_tn(pb_ctx* ctx, const double* W, int64_t ldw, int64_t n, int64_t rows, int splits, int64_t rows_per_split,
              double* out, int64_t ldo, int64_t split_stride);
namespace {
thread_local bool g_time_tn = false;   // set per call by pbi_gemm_tn; thread-local: several contexts may run in parallel threads (multi.cu)

struct Batch { int count = 1; int64_t sA = 0, sB = 0, sC = 0; };

template <int BM, int BN, int WARPS_M, int WARPS_N>
int launch_tn(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka, const double* B, int64_t ldb,
              int64_t kb, int64_t rows, double* C, int64_t ldc, int sym, Batch bt = Batch()) {
    constexpr int STAGES = 4;
    const int tiles_m = (int)((ka + BM - 1) / BM), tiles_n = (int)((kb + BN - 1) / BN);
    const int tiles = sym ? tiles_n * (tiles_n + 1) / 2 : tiles_m * tiles_n;
    const int splits = choose_splits(tiles * bt.count, rows, ctx->sm_count * ((BM <= 32) ? 2 : 1));
    int64_t rps = (rows + splits - 1) / splits;
    rps = (rps + BK - 1) / BK * BK;
    const bool direct = (splits == 1 && !sym);
    const int64_t split_stride = (int64_t)ka * kb;
    double* out = C; int64_t ldo = ldc;
    if (!direct) {
        PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)split_stride * splits * bt.count));
        out = ctx->ws; ldo = kb;
    }
    const int64_t bsOut = direct ? bt.sC : split_stride * splits;
    const bool al = aligned16(A, lda) && aligned16(B, ldb) && (bt.sA % 2 == 0) && (bt.sB % 2 == 0);
    size_t smem = (size_t)STAGES * BK * ((BM + 4) + (BN + 4)) * sizeof(double);
    dim3 grid(tiles, splits, bt.count);
    if (sym && BM == 128 && BN == 128 && bt.count == 1) {     // stream-K TMA kernel: writes C itself
        int rc = pbi_gram_streamk(ctx, A, lda, ka, rows, C, ldc, g_time_tn);
        if (rc < 0) return rc;
        if (rc == PB_OK) return PB_OK;                         // its reduction mirrors the lower triangle itself
    }
    if (sym && BM == 128 && BN == 128 && bt.count == 1) {     // TMA + mbarrier pipeline when the layout allows it
        int rc = pbi_gram_tma(ctx, A, lda, ka, rows, splits, rps, ctx->ws, split_stride, g_time_tn);
        if (rc < 0) return rc;
        if (rc == PB_OK) goto reduce;
    }
    if (!sym && BM == 64 && BN == 256 && bt.count == 1 && A == B && lda == ldb && ka == 64) {
        // the QR trailing product Y^T [Y | C]: TMA + mbarrier kernel (qr_tn.cu), same split layout
        int rc = pbi_qr_tn(ctx, A, lda, kb, rows, splits, rps, out, ldo, split_stride);
        if (rc < 0) return rc;
        if (rc == PB_OK) goto reduce;
    }
    if (g_time_tn) cudaEventRecord(ctx->evk0, ctx->stream);
    if (al) {
        auto kern = gemm_tn_kernel<BM, BN, WARPS_M, WARPS_N, STAGES, true>;
        PB_TRY(set_smem(ctx, kern, smem));
        kern<<<grid, NTHREADS, smem, ctx->stream>>>(A, lda, (int)ka, B, ldb, (int)kb, rows, rps, out, ldo,
                                                    split_stride, tiles_m, tiles_n, sym, bt.sA, bt.sB, bsOut);
    } else {
        auto kern = gemm_tn_kernel<BM, BN, WARPS_M, WARPS_N, STAGES, false>;
        PB_TRY(set_smem(ctx, kern, smem));
        kern<<<grid, NTHREADS, smem, ctx->stream>>>(A, lda, (int)ka, B, ldb, (int)kb, rows, rps, out, ldo,
                                                    split_stride, tiles_m, tiles_n, sym, bt.sA, bt.sB, bsOut);
    }
    PB_LAUNCH_CHECK(ctx);
    if (g_time_tn) { cudaEventRecord(ctx->evk1, ctx->stream); ctx->tn_timing_pending = true; }
reduce:
    if (!direct) {
        int64_t n = (int64_t)ka * kb;
        reduce_splits_kernel<<<dim3((unsigned)((n + 255) / 256), bt.count), 256, 0, ctx->stream>>>(
            ctx->ws, splits, split_stride, (int)ka, (int)kb, kb, C, ldc, sym, BM, BN, split_stride * splits, bt.sC);
        PB_LAUNCH_CHECK(ctx);
        if (sym) {
            symmetrize_kernel<<<dim3((unsigned)((n + 255) / 256), bt.count), 256, 0, ctx->stream>>>(C, (int)ka, ldc, bt.sC);
            PB_LAUNCH_CHECK(ctx);
        }
    }
    return PB_OK;
}

template <int BN, int WARPS_M, int WARPS_N, int STAGES = 4>
int launch_nn(pb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t m, const double* B,
              int64_t ldb, int64_t k, double* Y, int64_t ldy, const double* Uref, int64_t ldu,
              double* sig_part, Batch bt = Batch(), int subtract = 0) {
    size_t smem = (size_t)STAGES * (128 * (BK + 4) + BK * (BN + 4)) * sizeof(double);
    dim3 grid((unsigned)((rows + 127) / 128), (unsigned)((k + BN - 1) / BN), bt.count);
    const bool aa = aligned16(A, lda) && (bt.sA % 2 == 0), ab = aligned16(B, ldb) && (bt.sB % 2 == 0);
#define PB_NN_CASE(AA, AB) { auto kern = gemm_nn_kernel<BN, WARPS_M, WARPS_N, STAGES, AA, AB>; \
        PB_TRY(set_smem(ctx, kern, smem)); \
        kern<<<grid, NTHREADS, smem, ctx->stream>>>(A, lda, rows, (int)m, B, ldb, (int)k, Y, ldy, Uref, ldu, sig_part, \
                                                    bt.sA, bt.sB, bt.sC, subtract); }
    if (aa && ab) PB_NN_CASE(true, true)
    else if (aa) PB_NN_CASE(true, false)
    else if (ab) PB_NN_CASE(false, true)
    else PB_NN_CASE(false, false)
#undef PB_NN_CASE
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}


template <int KP, int MS = 32>
int launch_recon_strip(pb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t m, const double* B,
                       int64_t ldb, int64_t k, double* Y, int64_t ldy, const double* Uref, int64_t ldu, double* sig_out,
                       int subtract = 0) {
    constexpr int STG = (KP <= 16) ? 3 : 2;
    constexpr size_t smem = (size_t)(MS * (KP + 4) + STG * KP * 68 + STG * MS * 72) * sizeof(double);
    const bool aa = aligned16(A, lda), ab = aligned16(B, ldb), au = !Uref || aligned16(Uref, ldu);
    dim3 grid((unsigned)((rows + MS - 1) / MS));
#define PB_RS_CASE(AA, AB, AU) { auto kern = recon_strip_kernel<KP, MS, AA, AB, AU>; \
        PB_TRY(set_smem(ctx, kern, smem)); \
        kern<<<grid, NTHREADS, smem, ctx->stream>>>(A, lda, rows, (int)m, B, ldb, k, Y, ldy, Uref, ldu, sig_out, subtract); }
    if (aa && ab && au) PB_RS_CASE(true, true, true)
    else if (ab && au) PB_RS_CASE(false, true, true)
    else if (ab) PB_RS_CASE(false, true, false)
    else PB_RS_CASE(false, false, false)
#undef PB_RS_CASE
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

}  // namespace

int pbi_gemm_tn(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka, const double* B, int64_t ldb,
                int64_t kb, int64_t rows, double* C, int64_t ldc, int flags) {
    const int sym = flags & 1;
    g_time_tn = (flags & 2) != 0;
    if (ka <= 0 || kb <= 0 || rows < 0) PB_FAIL(ctx, PB_ERR_SHAPE, "gemm_tn: bad shape ka=%lld kb=%lld rows=%lld",
                                                (long long)ka, (long long)kb, (long long)rows);
    if (ka > (1 << 20) || kb > (1 << 20)) PB_FAIL(ctx, PB_ERR_SHAPE, "gemm_tn: ka/kb too large");
    if (sym && (A != B || ka != kb || lda != ldb)) PB_FAIL(ctx, PB_ERR_ARG, "gemm_tn: sym needs A == B");
    if (rows == 0) {
        PB_CUDA(ctx, cudaMemset2DAsync(C, ldc * sizeof(double), 0, kb * sizeof(double), ka, ctx->stream));
        return PB_OK;
    }
    if (ka <= 32 && kb <= 32) {      // skinny x skinny (second-pass Gram of the rotated block): HBM bound
        PB_TRY((launch_tn<32, 32, 2, 4>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0)));
        if (sym) { symmetrize_kernel<<<(unsigned)((ka * kb + 255) / 256), 256, 0, ctx->stream>>>(C, (int)ka, ldc, 0); PB_LAUNCH_CHECK(ctx); }
        return PB_OK;
    }
    if (!sym && ka > 32 && ka <= 64 && kb <= 64) return launch_tn<64, 64, 2, 4>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0);   // Y^T Y of a QR panel
    {   // wide right operand (the QR trailing product [Y^T Y | Y^T C]): 64 x 256 tiles, warp tile 32 x 64 (12 fragment
        // loads per 32 DMMAs instead of 8 per 16)
        static const int wide_env = getenv("PB_TN_WIDE") ? atoi(getenv("PB_TN_WIDE")) : 1;
        if (wide_env && !sym && ka > 32 && ka <= 64 && kb >= 512)
            return launch_tn<64, 256, 2, 4>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0);
    }
    if (!sym && ka > 32 && ka <= 64) return launch_tn<64, 128, 2, 4>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0);
    // 64 < ka <= 96 (the projection V^T U with n_L = 67 at C5): tiles as narrow as ka rounded up to 8 -- the DMMA count
    // follows the padded width (a 96-wide tile for 67 columns wastes 30 % of the pipe; 72-wide: 7 %)
    if (!sym && ka > 64 && ka <= 72) return launch_tn<72, 128, 1, 8>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0);
    if (!sym && ka > 72 && ka <= 80) return launch_tn<80, 128, 1, 8>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0);
    if (!sym && ka > 80 && ka <= 88) return launch_tn<88, 128, 1, 8>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0);
    if (!sym && ka > 88 && ka <= 96) return launch_tn<96, 128, 2, 4>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0);
    if (sym || ka > 32) return launch_tn<128, 128, 2, 4>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, sym);
    if (ka <= 16) return launch_tn<16, 128, 1, 8>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0);   // n_L <= 16: HBM bound
    return launch_tn<32, 128, 1, 8>(ctx, A, lda, ka, B, ldb, kb, rows, C, ldc, 0);
}

int pbi_gemm_nn(pb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t m, const double* B,
                int64_t ldb, int64_t k, double* Y, int64_t ldy, const double* Uref, int64_t ldu,
                double* sig_out) {
    if (rows <= 0 || m <= 0 || k <= 0) PB_FAIL(ctx, PB_ERR_SHAPE, "gemm_nn: bad shape rows=%lld m=%lld k=%lld",
                                               (long long)rows, (long long)m, (long long)k);
    const bool moments_mode = Uref && ldu < 0;
    {   // reconstruction / pod_sig with n_L <= 64 and a wide output: the strip kernel (HBM streaming)
        static const int strip_env = getenv("PB_NN_STRIP") ? atoi(getenv("PB_NN_STRIP")) : 1;
        if (strip_env && !moments_mode && m <= 64 && k >= 256) {
            if (m <= 16) return launch_recon_strip<16>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, sig_out);
            if (m <= 32) return launch_recon_strip<32>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, sig_out);
            return launch_recon_strip<64>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, sig_out);
        }
    }
    double* part = nullptr;
    const bool wide = k > 32;
    const int bn = k > 96 ? 128 : k > 64 ? 96 : k > 32 ? 64 : (k <= 16 ? 16 : 32);
    const int ntiles = (int)((k + bn - 1) / bn);
    const bool moments = Uref && ldu < 0;
    if (moments) part = sig_out;                     // running sum of squares, accumulated in place
    else if (Uref) {
        PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)ntiles * rows));
        part = ctx->ws;
    }
    if (k > 96) PB_TRY((launch_nn<128, 2, 4>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, part)));
    else if (k > 88) PB_TRY((launch_nn<96, 2, 4>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, part)));
    else if (k > 80) PB_TRY((launch_nn<88, 8, 1>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, part)));
    else if (k > 72) PB_TRY((launch_nn<80, 8, 1>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, part)));
    else if (k > 64) PB_TRY((launch_nn<72, 8, 1>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, part)));
    else if (wide) PB_TRY((launch_nn<64, 4, 2>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, part)));
    else if (k <= 16) PB_TRY((launch_nn<16, 8, 1>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, part)));
    else      PB_TRY((launch_nn<32, 8, 1>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Uref, ldu, part)));
    if (Uref && !moments) {
        sig_reduce_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(
            part, ntiles, rows, 0.5 / (double)k, sig_out);
        PB_LAUNCH_CHECK(ctx);
    }
    return PB_OK;
}

// Y += A * B in place (the QR trailing update passes B = -X)
int pbi_gemm_nn_sub(pb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t m, const double* B, int64_t ldb,
                    int64_t k, double* Y, int64_t ldy) {
    if (rows <= 0 || m <= 0 || k <= 0) PB_FAIL(ctx, PB_ERR_SHAPE, "gemm_nn_sub: bad shape");
    if (m == 64) {      // full 64-column panels: the TMA / mbarrier / register-resident-C kernel (qr_update.cu)
        const int rc = pbi_qr_update(ctx, A, lda, rows, m, B, ldb, k, Y, ldy);
        if (rc <= 0) return rc;
    }
    // K <= 64 and a wide C: the strip kernel streams C through shared memory once (in place)
    static const int strip_env = getenv("PB_QR_STRIP") ? atoi(getenv("PB_QR_STRIP")) : 1;
    if (strip_env && m <= 64 && k >= 64) {
        if (m <= 16) return launch_recon_strip<16>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Y, ldy, nullptr, 1);
        if (m <= 32) return launch_recon_strip<32>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Y, ldy, nullptr, 1);
        static const int ms64 = getenv("PB_QR_STRIP_ROWS") ? atoi(getenv("PB_QR_STRIP_ROWS")) : 64;
        if (ms64 == 64) return launch_recon_strip<64, 64>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Y, ldy, nullptr, 1);
        return launch_recon_strip<64>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, Y, ldy, nullptr, 1);
    }
    if (k > 96) return launch_nn<128, 2, 4>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, nullptr, 0, nullptr, Batch(), 1);
    if (k > 64) return launch_nn<96, 2, 4>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, nullptr, 0, nullptr, Batch(), 1);
    if (k > 32) return launch_nn<64, 4, 2>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, nullptr, 0, nullptr, Batch(), 1);
    if (k > 16) return launch_nn<32, 8, 1>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, nullptr, 0, nullptr, Batch(), 1);
    return launch_nn<16, 8, 1>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, nullptr, 0, nullptr, Batch(), 1);
}

int pbi_transpose(pb_ctx* ctx, const double* in, int64_t rows, int64_t cols, int64_t ldin,
                  double* out, int64_t ldout) {
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, ctx->stream>>>(in, rows, cols, ldin, out, ldout);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

// batch of `count` independent products: A_z = A + z*sA (same lda), C_z = C + z*sC  (two-step POD)
int pbi_gram_batched(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka, int64_t sA, int64_t rows,
                     double* C, int64_t sC, int count) {
    Batch bt; bt.count = count; bt.sA = sA; bt.sB = sA; bt.sC = sC;
    g_time_tn = false;
    return launch_tn<128, 128, 2, 4>(ctx, A, lda, ka, A, lda, ka, rows, C, ka, 1, bt);
}
int pbi_gemm_nn_batched(pb_ctx* ctx, const double* A, int64_t lda, int64_t sA, int64_t rows, int64_t m,
                        const double* B, int64_t ldb, int64_t sB, int64_t k, double* Y, int64_t ldy, int64_t sY, int count) {
    Batch bt; bt.count = count; bt.sA = sA; bt.sB = sB; bt.sC = sY;
    if (k > 32) return launch_nn<128, 2, 4>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, nullptr, 0, nullptr, bt);
    return launch_nn<32, 8, 1>(ctx, A, lda, rows, m, B, ldb, k, Y, ldy, nullptr, 0, nullptr, bt);
}
