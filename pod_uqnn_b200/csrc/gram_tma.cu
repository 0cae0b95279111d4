// Gram G = A^T A (SYRK-shaped, upper tiles) with a TMA + mbarrier warp-specialised pipeline.
//
// Same tiling and DMMA fragment mapping as gemm_tn_kernel<128,128> (gemm.cu), but the operand
// panels are brought in by the Tensor Memory Accelerator (cp.async.bulk.tensor, SASS UTMALDG)
// issued by one producer thread, and the 8 consumer warps never meet at a CTA-wide barrier:
// a stage is handed over with "full" / "empty" mbarriers, so warps drift and the DMMA pipe stays
// fed across stage boundaries (the cp.async version loses ~6 % of its issue samples at
// __syncthreads and spends issue slots on 8 LDGSTS + address math per thread per stage).
//
// Shared-memory layout of a panel (128 columns x 16 rows of A): the tensor map views the row-major
// matrix as (8 columns, rows, column groups) and the box (8, 16, 16) lands as [group][row][8 cols]
// with the 64-byte swizzle, i.e. the 16-byte chunk index is XORed with (row >> 1) & 3.  A DMMA
// k-step takes rows {r, r+1, r+4, r+5} of an 8-row block (any 4 rows may form a k-step: the sum
// over k is order independent), which makes the 16 lanes of a half-warp hit 16 distinct 8-byte
// banks -- no padding, no conflicts.
#include "common.cuh"
#include "tile_ops.cuh"
#include <cuda.h>
#include <cstdlib>
#include <vector>
#include <algorithm>

namespace {
using pbt::dmma884;

constexpr int TBK = 16;                 // rows per stage
constexpr int TSTAGES = 6;
constexpr int PANEL_DOUBLES = 128 * TBK;            // 2048 doubles = 16 KB
constexpr int STAGE_BYTES = 2 * PANEL_DOUBLES * 8;  // A panel + B panel
constexpr int TMA_THREADS = 288;        // 8 consumer warps + 1 producer warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
gram_tma_kernel(const __grid_constant__ CUtensorMap tmap, int ka, int64_t rows, int64_t rows_per_split,
                double* __restrict__ out, int64_t ldo, int64_t split_stride, int tiles_n) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    // the swizzle is a function of the absolute shared address: align the panels to 1 KB
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    double* panels = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TSTAGES * STAGE_BYTES);
    uint64_t* empty = full + TSTAGES;

    int ti = 0, rem = blockIdx.x;            // linear id -> (ti, tj >= ti)
    while (rem >= tiles_n - ti) { rem -= tiles_n - ti; ++ti; }
    const int tj = ti + rem;
    const bool diag = (ti == tj);
    const int split = blockIdx.y;
    const int64_t r_begin = (int64_t)split * rows_per_split;
    const int64_t r_end = min(rows, r_begin + rows_per_split);
    const int nk = r_end > r_begin ? (int)((r_end - r_begin + TBK - 1) / TBK) : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < TSTAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        // ---------------- producer: one thread issues the TMA loads
        if (lane == 0) {
            for (int k = 0; k < nk; ++k) {
                const int s = k % TSTAGES;
                const uint32_t ph = (k / TSTAGES) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_expect_tx(full + s, diag ? STAGE_BYTES / 2 : STAGE_BYTES);
                double* dst = panels + (size_t)s * 2 * PANEL_DOUBLES;
                const int r0 = (int)(r_begin + (int64_t)k * TBK);
                tma_load_3d(dst, &tmap, full + s, 0, r0, ti * 16);
                if (!diag) tma_load_3d(dst + PANEL_DOUBLES, &tmap, full + s, 0, r0, tj * 16);
            }
        }
        return;
    }

    // ---------------- consumers: 8 warps as 2 x 4, warp tile 64 x 32
    const int wm = warp >> 2, wn = warp & 3;
    const int lk = lane & 3, lg = lane >> 2;
    double acc[8][4][2];
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }
    // offset (in doubles, inside a [16 rows][8 cols] group block) of this lane's element for k-step kk
    int koff[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int k = (kk >> 1) * 8 + (kk & 1) * 2 + (lk & 1) + (lk >> 1) * 4;
        koff[kk] = k * 8 + ((((lg >> 1) ^ ((k >> 1) & 3))) << 1) + (lg & 1);
    }
    for (int k = 0; k < nk; ++k) {
        const int s = k % TSTAGES;
        const uint32_t ph = (k / TSTAGES) & 1;
        mbar_wait(full + s, ph);
        const double* As = panels + (size_t)s * 2 * PANEL_DOUBLES;
        const double* Bs = diag ? As : As + PANEL_DOUBLES;
        const double* ap = As + (wm * 8) * (TBK * 8);
        const double* bp = Bs + (wn * 4) * (TBK * 8);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int t = 0; t < 8; ++t) a[t] = ap[t * (TBK * 8) + koff[kk]];
#pragma unroll
            for (int u = 0; u < 4; ++u) b[u] = bp[u * (TBK * 8) + koff[kk]];
#pragma unroll
            for (int t = 0; t < 8; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }

    const int64_t i0 = (int64_t)ti * 128, j0 = (int64_t)tj * 128;
    double* o = out + (int64_t)split * split_stride;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int64_t i = i0 + wm * 64 + t * 8 + lg;
        if (i >= ka) continue;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = j0 + wn * 32 + u * 8 + 2 * lk;
            if (j < ka)     o[i * ldo + j]     = acc[t][u][0];
            if (j + 1 < ka) o[i * ldo + j + 1] = acc[t][u][1];
        }
    }
}

// ---------------------------------------------------------------------------
// stream-K variant: one persistent CTA per SM owns a contiguous range of the linearised
// (tile, k-iteration) space, weighted 4 per off-diagonal and 3 per diagonal iteration.
//   * no wave quantisation and no split heuristics: every SM gets the same weight;
//   * diagonal tiles skip their strictly-lower 64 x 64 quadrant: warps 0-3 keep their 64 x 32 tiles of
//     the upper half, warps 4-7 take 64 x 16 tiles of the lower-right quadrant, so each SM sub-partition
//     issues 3/4 of the DMMAs of an off-diagonal tile (the consumer warps are decoupled, no barrier);
//   * the partition is static (host tables), so the partial tiles are summed in a fixed order.
struct SegEntry { int ti, tj, k0, k1, slot; };
struct GramPlan { int64_t ka = -1, rows = -1; int P = 0; int nslots = 0; std::vector<SegEntry> segs; std::vector<int> seg_begin,
                  tile_off, tile_src; int* d_tab = nullptr; size_t d_cap = 0; };

__global__ void __launch_bounds__(TMA_THREADS, 1)
gram_streamk_kernel(const __grid_constant__ CUtensorMap tmap, const SegEntry* __restrict__ segs,
                    const int* __restrict__ seg_begin, double* __restrict__ ws) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    double* panels = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TSTAGES * STAGE_BYTES);
    uint64_t* empty = full + TSTAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < TSTAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    const int sb = seg_begin[blockIdx.x], se = seg_begin[blockIdx.x + 1];

    if (warp == 8) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int sg = sb; sg < se; ++sg) {
                const SegEntry e = segs[sg];
                const bool diag = (e.ti == e.tj);
                for (int k = e.k0; k < e.k1; ++k, ++it) {
                    const int s = it % TSTAGES; const uint32_t ph = (it / TSTAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_expect_tx(full + s, diag ? STAGE_BYTES / 2 : STAGE_BYTES);
                    double* dst = panels + (size_t)s * 2 * PANEL_DOUBLES;
                    tma_load_3d(dst, &tmap, full + s, 0, k * TBK, e.ti * 16);
                    if (!diag) tma_load_3d(dst + PANEL_DOUBLES, &tmap, full + s, 0, k * TBK, e.tj * 16);
                }
            }
        }
        return;
    }

    const int wm = warp >> 2, wn = warp & 3;
    const int lk = lane & 3, lg = lane >> 2;
    int koff[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int k = (kk >> 1) * 8 + (kk & 1) * 2 + (lk & 1) + (lk >> 1) * 4;
        koff[kk] = k * 8 + ((((lg >> 1) ^ ((k >> 1) & 3))) << 1) + (lg & 1);
    }
    uint32_t it = 0;
    for (int sg = sb; sg < se; ++sg) {
        const SegEntry e = segs[sg];
        const bool diag = (e.ti == e.tj);
        const bool half = diag && wm == 1;                 // lower-right quadrant only, 16 columns per warp
        const int col0 = half ? 64 + wn * 16 : wn * 32;   // first column of this warp inside the tile
        double acc[8][4][2];
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }
        for (int k = e.k0; k < e.k1; ++k, ++it) {
            const int s = it % TSTAGES; const uint32_t ph = (it / TSTAGES) & 1;
            mbar_wait(full + s, ph);
            const double* As = panels + (size_t)s * 2 * PANEL_DOUBLES;
            const double* Bs = diag ? As : As + PANEL_DOUBLES;
            const double* ap = As + (wm * 8) * (TBK * 8);
            const double* bp = Bs + (col0 / 8) * (TBK * 8);
            if (!half) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    double a[8], b[4];
#pragma unroll
                    for (int t = 0; t < 8; ++t) a[t] = ap[t * (TBK * 8) + koff[kk]];
#pragma unroll
                    for (int u = 0; u < 4; ++u) b[u] = bp[u * (TBK * 8) + koff[kk]];
#pragma unroll
                    for (int t = 0; t < 8; ++t)
#pragma unroll
                        for (int u = 0; u < 4; ++u) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    double a[8], b[2];
#pragma unroll
                    for (int t = 0; t < 8; ++t) a[t] = ap[t * (TBK * 8) + koff[kk]];
#pragma unroll
                    for (int u = 0; u < 2; ++u) b[u] = bp[u * (TBK * 8) + koff[kk]];
#pragma unroll
                    for (int t = 0; t < 8; ++t)
#pragma unroll
                        for (int u = 0; u < 2; ++u) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        // partial tile -> workspace slot (dense 128 x 128)
        double* o = ws + (size_t)e.slot * (128 * 128);
        const int nu = half ? 2 : 4;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int i = wm * 64 + t * 8 + lg;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u < nu) {
                    const int j = col0 + u * 8 + 2 * lk;
                    *reinterpret_cast<double2*>(o + i * 128 + j) = make_double2(acc[t][u][0], acc[t][u][1]);
                }
            }
        }
    }
}

// C[i][j] = sum over the slots of tile (ti, tj) in table order; lower tiles mirrored from the upper ones
__global__ void streamk_reduce_kernel(const double* __restrict__ ws, const int* __restrict__ tile_src_off,
                                      const int* __restrict__ tile_src, int ka, int tiles_n,
                                      double* __restrict__ C, int64_t ldc) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)ka * ka) return;
    const int i = (int)(idx / ka), j = (int)(idx % ka);
    int si = i, sj = j;
    if (j / 128 < i / 128) { si = j; sj = i; }
    const int ti = si / 128, tj = sj / 128;
    const int t = ti * tiles_n - ti * (ti - 1) / 2 + (tj - ti);
    const int li = si % 128, lj = sj % 128;
    double s = 0.;
    for (int q = tile_src_off[t]; q < tile_src_off[t + 1]; ++q) s += ws[(size_t)tile_src[q] * (128 * 128) + li * 128 + lj];
    C[(int64_t)i * ldc + j] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static const EncodeTiledFn fn = []() -> EncodeTiledFn {      // initialised once, thread-safe (several contexts may call at once)
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
        return nullptr;
    }();
    return fn;
}

}  // namespace

// 3-D fp64 tensor map for the other TMA kernels of the library (qr_update.cu); 0 on success
int pbi_tma_encode_3d(CUtensorMap* map, const double* base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                      const uint32_t box[3], int swizzle64) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return 1;
    const cuuint64_t gdim[3] = {dims[0], dims[1], dims[2]};
    const cuuint64_t gstr[2] = {strides_bytes[0], strides_bytes[1]};
    const cuuint32_t bx[3] = {box[0], box[1], box[2]};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstr, bx, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

void pbi_gram_plan_free(pb_ctx* ctx) {
    if (!ctx->gram_plan) return;
    GramPlan* p = static_cast<GramPlan*>(ctx->gram_plan);
    if (p->d_tab) cudaFree(p->d_tab);
    delete p; ctx->gram_plan = nullptr;
}

// returns PB_OK when the TMA kernel was launched, 1 when the caller should use the cp.async kernel
int pbi_gram_tma(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka, int64_t rows, int splits,
                 int64_t rows_per_split, double* ws, int64_t split_stride, bool time_it) {
    static const bool enabled = []{ const char* e = getenv("PB_GRAM_TMA"); return !e || atoi(e) != 0; }();
    if (!enabled) return 1;
    if ((lda % 2) || (ka % 8) || (reinterpret_cast<uintptr_t>(A) % 16) || rows >= (1ll << 31) || rows_per_split % TBK) return 1;
    EncodeTiledFn enc = get_encode();
    if (!enc) return 1;
    CUtensorMap map;
    const cuuint64_t gdim[3] = {8, (cuuint64_t)rows, (cuuint64_t)(ka / 8)};
    const cuuint64_t gstr[2] = {(cuuint64_t)lda * 8, 64};        // bytes: rows, column groups
    const cuuint32_t box[3] = {8, (cuuint32_t)TBK, 16};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(A), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1;
    const int tiles_n = (int)((ka + 127) / 128);
    const int tiles = tiles_n * (tiles_n + 1) / 2;
    const size_t smem = (size_t)TSTAGES * STAGE_BYTES + 2 * TSTAGES * sizeof(uint64_t) + 1024;
    PB_TRY(pbi_smem_optin(ctx, (const void*)gram_tma_kernel, (size_t)(smem)));
    if (time_it) cudaEventRecord(ctx->evk0, ctx->stream);
    gram_tma_kernel<<<dim3(tiles, splits), TMA_THREADS, smem, ctx->stream>>>(map, (int)ka, rows, rows_per_split, ws, ka,
                                                                            split_stride, tiles_n);
    PB_LAUNCH_CHECK(ctx);
    if (time_it) { cudaEventRecord(ctx->evk1, ctx->stream); ctx->tn_timing_pending = true; }
    return PB_OK;
}


// stream-K launcher: returns PB_OK when it produced C (upper part; the caller symmetrises), 1 to fall back
int pbi_gram_streamk(pb_ctx* ctx, const double* A, int64_t lda, int64_t ka, int64_t rows, double* C, int64_t ldc,
                     bool time_it) {
    static const bool enabled = []{ const char* e = getenv("PB_GRAM_STREAMK"); return !e || atoi(e) != 0; }();
    if (!enabled) return 1;
    if ((lda % 2) || (ka % 8) || (reinterpret_cast<uintptr_t>(A) % 16) || rows >= (1ll << 31) - 64) return 1;
    EncodeTiledFn enc = get_encode();
    if (!enc) return 1;
    const int tiles_n = (int)((ka + 127) / 128);
    const int T = tiles_n * (tiles_n + 1) / 2;
    const int P = ctx->sm_count;
    const int64_t nk = (rows + TBK - 1) / TBK;
    if (nk * T < 4 * (int64_t)P) return 1;                 // too little work to spread: use the tiled kernel
    CUtensorMap map;
    const cuuint64_t gdim[3] = {8, (cuuint64_t)rows, (cuuint64_t)(ka / 8)};
    const cuuint64_t gstr[2] = {(cuuint64_t)lda * 8, 64};
    const cuuint32_t box[3] = {8, (cuuint32_t)TBK, 16};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(A), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;

    // ---- static partition (cached per shape)
    // the schedule and its device table belong to the context (one GPU each): another context in the same thread builds its own
    if (!ctx->gram_plan) ctx->gram_plan = new GramPlan();
    GramPlan& plan = *static_cast<GramPlan*>(ctx->gram_plan);
    if (plan.ka != ka || plan.rows != rows || plan.P != P) {
        plan.ka = ka; plan.rows = rows; plan.P = P; plan.segs.clear(); plan.seg_begin.assign(P + 1, 0);
        // Static schedule.  Two parts:
        //  (1) ALIGNED part: equal-length pieces of OFF-DIAGONAL tiles that all start together, one per CTA
        //      and round -- pieces with the same row range run in lockstep on different SMs, so a panel
        //      (128 columns x that row range) is fetched from HBM once and served to the ~tiles_n tiles
        //      that need it from L2.  T < P: a tile is cut into q pieces of L = floor(D/4) iterations
        //      (D = total weight / P); T >= P: the pieces are whole tiles, floor(n_off / P) rounds.
        //  (2) TAIL: what is left (piece remainders, surplus tiles, all diagonal tiles) as one weighted
        //      linear space, split so that every CTA ends with the same total weight D (stream-K).
        // A plain contiguous stream-K split measured 10.5 GB of DRAM reads for the 1.28 GB matrix at C2
        // (ranges start at unrelated rows, nothing is shared); this schedule keeps the traffic near the
        // algorithmic size and the DMMA pipe at ~95 %.
        struct Item { int ti, tj; int64_t k0, k1; };
        std::vector<Item> tail;
        std::vector<std::vector<Item>> own(P);
        std::vector<std::pair<int,int>> off, dia;
        for (int i = 0; i < tiles_n; ++i) for (int j = i; j < tiles_n; ++j) (i == j ? dia : off).push_back({i, j});
        const int n_off = (int)off.size();
        const double Wtot = (double)nk * (4. * n_off + 3. * dia.size());
        const double D = Wtot / P;
        if (n_off > 0 && 4. * nk >= D) {                       // pieces of tiles, one round
            const int64_t L = std::max<int64_t>(1, (int64_t)(D / 4.));
            const int64_t q = nk / L;
            int cta = 0;
            bool diag_aligned = false;
            for (int64_t pc = 0; pc < q; ++pc)                 // piece-major: CTAs of a row range are neighbours
                for (int t = 0; t < n_off; ++t) {
                    Item it{off[t].first, off[t].second, pc * L, (pc + 1) * L};
                    if (cta < P) own[cta++].push_back(it); else tail.push_back(it);
                }
            // diagonal tiles in the same row ranges (they cost 3/4 and top up from the tail) when CTAs are left
            if (cta + (int)dia.size() * q <= P) {
                diag_aligned = true;
                for (int64_t pc = 0; pc < q; ++pc)
                    for (auto& d : dia) own[cta++].push_back({d.first, d.second, pc * L, (pc + 1) * L});
            }
            if (q * L < nk) {
                for (int t = 0; t < n_off; ++t) tail.push_back({off[t].first, off[t].second, q * L, nk});
                if (diag_aligned) for (auto& d : dia) tail.push_back({d.first, d.second, q * L, nk});
            }
            if (diag_aligned) dia.clear();
        } else {                                               // whole tiles, several rounds
            const int rounds = n_off / P;
            for (int r = 0; r < rounds; ++r)
                for (int sI = 0; sI < P; ++sI) own[sI].push_back({off[r * P + sI].first, off[r * P + sI].second, 0, nk});
            for (int t = rounds * P; t < n_off; ++t) tail.push_back({off[t].first, off[t].second, 0, nk});
        }
        for (auto& d : dia) tail.push_back({d.first, d.second, 0, nk});
        // weights
        auto wof = [](const Item& it) { return (it.k1 - it.k0) * (it.ti == it.tj ? 3 : 4); };
        std::vector<int64_t> cum(tail.size() + 1, 0);
        for (size_t x = 0; x < tail.size(); ++x) cum[x + 1] = cum[x] + wof(tail[x]);
        const int64_t WT = cum.back();
        std::vector<int64_t> aw(P, 0); int64_t Wall = WT;
        for (int sI = 0; sI < P; ++sI) { for (auto& it : own[sI]) aw[sI] += wof(it); Wall += aw[sI]; }
        // tail share of CTA s: proportional to (Wall / P - aligned weight), >= 0
        std::vector<double> need(P); double need_sum = 0.;
        for (int sI = 0; sI < P; ++sI) { need[sI] = std::max(0., (double)Wall / P - (double)aw[sI]); need_sum += need[sI]; }
        auto locate = [&](int64_t b, int& x, int64_t& k) {
            if (b >= WT) { x = (int)tail.size(); k = 0; return; }
            int lo = (int)(std::upper_bound(cum.begin(), cum.end(), b) - cum.begin()) - 1;
            x = lo; k = (b - cum[lo]) / (tail[lo].ti == tail[lo].tj ? 3 : 4);
        };
        const int Tn = tiles_n;
        auto tile_id = [&](int ti, int tj) { return ti * Tn - ti * (ti - 1) / 2 + (tj - ti); };
        std::vector<std::vector<int>> src(T);
        int slot = 0;
        double acc_need = 0.;
        for (int sI = 0; sI < P; ++sI) {
            plan.seg_begin[sI] = (int)plan.segs.size();
            for (auto& it : own[sI]) {
                plan.segs.push_back({it.ti, it.tj, (int)it.k0, (int)it.k1, slot});
                src[tile_id(it.ti, it.tj)].push_back(slot); ++slot;
            }
            const int64_t b0 = need_sum > 0. ? (int64_t)((double)WT * (acc_need / need_sum)) : 0;
            acc_need += need[sI];
            const int64_t b1 = (sI + 1 == P) ? WT : (need_sum > 0. ? (int64_t)((double)WT * (acc_need / need_sum)) : 0);
            int x0, x1; int64_t k0, k1;
            locate(b0, x0, k0); locate(b1, x1, k1);
            for (int x = x0; x <= x1 && x < (int)tail.size(); ++x) {
                const int64_t ks = tail[x].k0 + (x == x0 ? k0 : 0);
                const int64_t ke = (x == x1) ? tail[x].k0 + k1 : tail[x].k1;
                if (ke > ks) {
                    plan.segs.push_back({tail[x].ti, tail[x].tj, (int)ks, (int)ke, slot});
                    src[tile_id(tail[x].ti, tail[x].tj)].push_back(slot); ++slot;
                }
            }
        }
        plan.seg_begin[P] = (int)plan.segs.size();
        plan.nslots = slot;
        plan.tile_off.assign(T + 1, 0); plan.tile_src.clear();
        for (int tt = 0; tt < T; ++tt) { plan.tile_off[tt] = (int)plan.tile_src.size(); for (int v : src[tt]) plan.tile_src.push_back(v); }
        plan.tile_off[T] = (int)plan.tile_src.size();
        // device tables: [segs (5 ints each)] [seg_begin (P+1)] [tile_off (T+1)] [tile_src]
        const size_t n_int = plan.segs.size() * 5 + (P + 1) + (T + 1) + plan.tile_src.size();
        if (plan.d_cap < n_int) { if (plan.d_tab) cudaFree(plan.d_tab); PB_CUDA(ctx, cudaMalloc(&plan.d_tab, n_int * sizeof(int))); plan.d_cap = n_int; }
        std::vector<int> host(n_int); size_t o = 0;
        memcpy(host.data(), plan.segs.data(), plan.segs.size() * 5 * sizeof(int)); o += plan.segs.size() * 5;
        memcpy(host.data() + o, plan.seg_begin.data(), (P + 1) * sizeof(int)); o += P + 1;
        memcpy(host.data() + o, plan.tile_off.data(), (T + 1) * sizeof(int)); o += T + 1;
        memcpy(host.data() + o, plan.tile_src.data(), plan.tile_src.size() * sizeof(int));
        PB_CUDA(ctx, cudaMemcpyAsync(plan.d_tab, host.data(), n_int * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const SegEntry* d_segs = reinterpret_cast<const SegEntry*>(plan.d_tab);
    const int* d_seg_begin = plan.d_tab + plan.segs.size() * 5;
    const int* d_tile_off = d_seg_begin + (P + 1);
    const int* d_tile_src = d_tile_off + (T + 1);
    PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)plan.nslots * 128 * 128));
    const size_t smem = (size_t)TSTAGES * STAGE_BYTES + 2 * TSTAGES * sizeof(uint64_t) + 1024;
    PB_TRY(pbi_smem_optin(ctx, (const void*)gram_streamk_kernel, (size_t)(smem)));
    if (time_it) cudaEventRecord(ctx->evk0, ctx->stream);
    gram_streamk_kernel<<<P, TMA_THREADS, smem, ctx->stream>>>(map, d_segs, d_seg_begin, ctx->ws);
    PB_LAUNCH_CHECK(ctx);
    if (time_it) { cudaEventRecord(ctx->evk1, ctx->stream); ctx->tn_timing_pending = true; }
    const int64_t n = ka * ka;
    streamk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->ws, d_tile_off, d_tile_src, (int)ka, tiles_n, C, ldc);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}
