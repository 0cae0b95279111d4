// Trailing product of the blocked Householder QR (qr.cu): Z (64 x n) = Y^T [Y | C], where Y = the 64 panel columns of the
// working copy W and [Y | C] = W[:, J:] -- the first 64-row tile of the Gram matrix of W[:, J:].  K = the rows of W
// (10^5 .. 10^6), so the product is split over row ranges and the partial tiles are summed by reduce_splits_kernel
// (gemm.cu), exactly like the cp.async kernel it replaces (gemm_tn_kernel<64,256>, DMMA pipe 78 %: every 16-row stage
// ended in a CTA barrier with the copy issue of all eight warps in the same window).
//
// Same recipe as the Gram kernel (gram_tma.cu): the row-major operand is viewed by ONE tensor map as (8 columns, rows,
// column groups); a box (8, 16, 8) = a 64-column x 16-row panel lands as [group][row][8 cols] with the 64-byte swizzle
// (cp.async.bulk.tensor.3d, SASS UTMALDG), a DMMA k-step takes rows {r, r+1, r+4, r+5} of an 8-row block, which keeps
// the 16 lanes of a half-warp on distinct banks without padding.  A stage = the Y panel + four panels of the 256-column
// B tile (40 KB), five stages handed over with full / empty mbarriers by one producer thread; the 8 consumer warps
// (2 x 4, warp tile 32 x 64: 12 fragment loads per 32 DMMAs) never meet at a CTA barrier.
#include "common.cuh"
#include "tile_ops.cuh"
#include <cuda.h>
#include <cstdlib>

int pbi_tma_encode_3d(CUtensorMap* map, const double* base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                      const uint32_t box[3], int swizzle64);

namespace {
using pbt::dmma884;

constexpr int ZBK = 16;                         // rows per stage
constexpr int ZSTAGES = 5;
constexpr int ZPANEL = 64 * ZBK;                // doubles of a 64-column panel (8 KB)
constexpr int ZSTAGE_DOUBLES = 5 * ZPANEL;      // Y panel + 4 B panels
constexpr int ZTHREADS = 288;                   // 8 consumer warps + 1 producer warp

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

// grid (column tiles of 256, row splits).  out: per split a 64 x n tile row (leading dimension ldo), element (i, j) =
// sum over the split's rows of W[h][i] W[h][j]
__global__ void __launch_bounds__(ZTHREADS, 1)
qr_tn_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n, int64_t rows, int64_t rows_per_split,
             double* __restrict__ out, int64_t ldo, int64_t split_stride) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // the swizzle follows the address
    double* panels = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)ZSTAGES * ZSTAGE_DOUBLES * 8);
    uint64_t* empty = full + ZSTAGES;

    const int tj = blockIdx.x, split = blockIdx.y;
    const int64_t r_begin = (int64_t)split * rows_per_split;
    const int64_t r_end = min(rows, r_begin + rows_per_split);
    const int nk = r_end > r_begin ? (int)((r_end - r_begin + ZBK - 1) / ZBK) : 0;
    // 64-column panels of this B tile that exist (the last tile of a row may be narrower)
    const int nbp = (int)min((int64_t)4, (n - (int64_t)tj * 256 + 63) / 64);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < ZSTAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        if (lane == 0) {
            for (int k = 0; k < nk; ++k) {
                const int s = k % ZSTAGES;
                const uint32_t ph = (k / ZSTAGES) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_expect_tx(full + s, (uint32_t)(1 + nbp) * ZPANEL * 8);
                double* dst = panels + (size_t)s * ZSTAGE_DOUBLES;
                const int r0 = (int)(r_begin + (int64_t)k * ZBK);
                tma_load_3d(dst, &tmap, full + s, 0, r0, 0);                                  // Y: columns [0, 64)
                for (int p = 0; p < nbp; ++p)
                    tma_load_3d(dst + (1 + p) * ZPANEL, &tmap, full + s, 0, r0, tj * 32 + p * 8);
            }
        }
        return;
    }

    // consumers: 2 x 4 warps, warp tile 32 rows (4 atoms) x 64 columns (8 atoms = one B panel)
    const int wm = warp >> 2, wn = warp & 3;
    const int lk = lane & 3, lg = lane >> 2;
    const bool active = wn < nbp;
    double acc[4][8][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }
    int koff[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int k = (kk >> 1) * 8 + (kk & 1) * 2 + (lk & 1) + (lk >> 1) * 4;
        koff[kk] = k * 8 + ((((lg >> 1) ^ ((k >> 1) & 3))) << 1) + (lg & 1);
    }
    for (int k = 0; k < nk; ++k) {
        const int s = k % ZSTAGES;
        const uint32_t ph = (k / ZSTAGES) & 1;
        mbar_wait(full + s, ph);
        if (active) {
            const double* ap = panels + (size_t)s * ZSTAGE_DOUBLES + (wm * 4) * (ZBK * 8);
            const double* bp = panels + (size_t)s * ZSTAGE_DOUBLES + (1 + wn) * ZPANEL;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                double a[4], b[8];
#pragma unroll
                for (int t = 0; t < 4; ++t) a[t] = ap[t * (ZBK * 8) + koff[kk]];
#pragma unroll
                for (int u = 0; u < 8; ++u) b[u] = bp[u * (ZBK * 8) + koff[kk]];
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 8; ++u) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    if (!active) return;
    double* o = out + (int64_t)split * split_stride;
    const int64_t j0 = (int64_t)tj * 256 + wn * 64 + 2 * lk;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int64_t i = wm * 32 + t * 8 + lg;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t j = j0 + u * 8;
            if (j < n) *reinterpret_cast<double2*>(o + i * ldo + j) = make_double2(acc[t][u][0], acc[t][u][1]);
        }
    }
}

}  // namespace

// returns PB_OK when the kernel was launched (the caller reduces the splits), 1 when the shape / alignment is outside
// its domain (the caller falls back to the cp.async kernel)
int pbi_qr_tn(pb_ctx* ctx, const double* W, int64_t ldw, int64_t n, int64_t rows, int splits, int64_t rows_per_split,
              double* out, int64_t ldo, int64_t split_stride) {
    static const bool enabled = []{ const char* e = getenv("PB_QR_TN_TMA"); return !e || atoi(e) != 0; }();
    if (!enabled || n < 64 || (n % 8) || (ldw % 2) || (ldo % 2) || (split_stride % 2) || rows >= (1ll << 31) - 64) return 1;
    if ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out)) % 16) return 1;
    if (rows_per_split % ZBK) return 1;
    CUtensorMap map;
    const uint64_t dims[3] = {8, (uint64_t)rows, (uint64_t)(n / 8)};
    const uint64_t strides[2] = {(uint64_t)ldw * 8, 64};
    const uint32_t box[3] = {8, (uint32_t)ZBK, 8};
    if (pbi_tma_encode_3d(&map, W, dims, strides, box, 1) != 0) return 1;
    const size_t smem = (size_t)ZSTAGES * ZSTAGE_DOUBLES * 8 + 2 * ZSTAGES * sizeof(uint64_t) + 1024;
    PB_TRY(pbi_smem_optin(ctx, (const void*)qr_tn_kernel, smem));
    const int tiles_n = (int)((n + 255) / 256);
    qr_tn_kernel<<<dim3(tiles_n, splits), ZTHREADS, smem, ctx->stream>>>(map, n, rows, rows_per_split, out, ldo, split_stride);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}
