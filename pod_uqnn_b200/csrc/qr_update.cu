// In-place trailing update of the blocked Householder QR (qr.cu): C (rows x n) += Y (rows x 64) * B (64 x n), B = -X.
// 2 rows x 64 x n flops per C element pair, C read and written once -> compute bound on the fp64 tensor pipe (16 flop per
// C byte); the cp.async strip kernel it replaces (gemm.cu recon_strip_kernel<64,64>) kept the DMMA pipe 57-62 % busy:
// every column tile ended in a CTA-wide barrier with the C tile staged through shared memory, so the epilogue stores,
// the accumulator refill and the copy issue of all eight warps fell into the same window.
//
// This kernel (persistent, one CTA per SM, 384 threads):
//   * (384 threads: the producer warpgroup gives its registers to the consumers, setmaxnreg 40 / 232)
//   * the CTA owns PAIRS of 64-row strips: consumer group g (4 warps, 2 x 2, warp tile 32 x 32) owns strip g.  The two
//     groups consume the SAME B tiles from one ring but never synchronise with each other, so they drift apart and one
//     group's epilogue / refill overlaps the other group's DMMA phase (every SM sub-partition hosts one warp of each);
//   * B tiles (64 x 64 doubles, 32 KB) arrive by ONE TMA instruction each (cp.async.bulk.tensor.3d, SASS UTMALDG)
//     issued by a producer thread into a 4-stage ring with full / empty mbarriers: the tensor map views the row-major
//     B as (8 columns, 64 rows, column groups), the box lands as [group][row][8 cols] with the 64-byte swizzle, and a
//     DMMA k-step takes rows {r, r+1, r+4, r+5} of an 8-row block (same bank-conflict-free scheme as gram_tma.cu);
//   * the C tile never passes through shared memory: each warp loads the NEXT tile's 32 x 32 block straight into a
//     second register set (16 LDG.128 in flight during the current tile's 256 DMMAs) and stores the finished one with
//     STG.128;
//   * the Y strip is staged once per strip (cp.async) with the k positions permuted to match the k-step order.
#include "common.cuh"
#include "tile_ops.cuh"
#include <cuda.h>
#include <cstdlib>
#include <algorithm>

int pbi_tma_encode_3d(CUtensorMap* map, const double* base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                      const uint32_t box[3], int swizzle64);

namespace {
using pbt::dmma884;
using pbt::cp_async16;
using pbt::cp_async_commit;

constexpr int KQ = 64;                  // panel width (reduction depth)
constexpr int NTQ = 64;                 // columns per B tile
constexpr int QSTAGES = 4;
constexpr int BTILE_DOUBLES = KQ * NTQ;             // 4096 doubles = 32 KB
constexpr int PAQ = KQ + 4;             // Y strip pitch (== 4 mod 16)
constexpr int QTHREADS = 384;           // 2 consumer groups x 4 warps + one producer warpgroup (registers re-balanced)

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
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;\n" :: "r"(1 + g) : "memory"); }

__global__ void __launch_bounds__(QTHREADS, 1)
qr_update_kernel(const __grid_constant__ CUtensorMap bmap, const double* __restrict__ Y, int64_t ldy, int64_t rows,
                 int64_t n, double* __restrict__ C, int64_t ldc, int n_pairs, int csplit) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    double* ring = reinterpret_cast<double*>(smem_raw);                         // [QSTAGES][8 groups][64 rows][8]
    double* As = ring + (size_t)QSTAGES * BTILE_DOUBLES;                        // [2 groups][64][PAQ]
    uint64_t* full = reinterpret_cast<uint64_t*>(As + 2 * 64 * PAQ);
    uint64_t* empty = full + QSTAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tiles = (int)((n + NTQ - 1) / NTQ);
    // work units: (strip pair, column range); csplit column ranges per pair even out the last wave of the persistent grid
    const int n_units = n_pairs * csplit, tiles_per = (n_tiles + csplit - 1) / csplit;
    if (tid == 0) {
        for (int s = 0; s < QSTAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    if (warp >= 8) {
        // ------------------------------------------------ producer: the B tiles, once per strip pair
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (warp == 8 && lane == 0) {
            uint32_t it = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const int t0 = (unit % csplit) * tiles_per, t1 = min(n_tiles, t0 + tiles_per);
                for (int t = t0; t < t1; ++t, ++it) {
                    const int s = it % QSTAGES; const uint32_t ph = (it / QSTAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_expect_tx(full + s, BTILE_DOUBLES * 8);
                    tma_load_3d(ring + (size_t)s * BTILE_DOUBLES, &bmap, full + s, 0, 0, t * (NTQ / 8));
                }
            }
        }
        return;
    }

    // ---------------------------------------------------- consumers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int g = warp >> 2, w = warp & 3;
    const int wm = w >> 1, wn = w & 1;
    const int lk = lane & 3, lg = lane >> 2;
    const int gt = tid & 127;                           // thread index inside the group
    double* as_g = As + g * 64 * PAQ;
    // offset (doubles, inside a [64 rows][8 cols] group block) of this lane's B element for k-step kk: rows {r, r+1,
    // r+4, r+5} of an 8-row block, 16-byte chunk XOR (row >> 1) & 3 (64-byte swizzle)
    int koff[16];
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
        const int k = (kk >> 1) * 8 + (kk & 1) * 2 + (lk & 1) + (lk >> 1) * 4;
        koff[kk] = k * 8 + ((((lg >> 1) ^ ((k >> 1) & 3))) << 1) + (lg & 1);
    }
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int pair = unit / csplit;
        const int t0 = (unit % csplit) * tiles_per, t1 = min(n_tiles, t0 + tiles_per);
        const int64_t h0 = ((int64_t)pair * 2 + g) * 64;
        const bool live = h0 < rows;                    // (the last pair may have only its first strip)
        // ---- Y strip: 64 rows x 64 k, 16-byte chunks, k positions permuted: chunk c of an 8-block goes to slot [0,2,1,3][c]
        group_sync(g);                                  // everyone in the group is done with the previous strip
        if (live) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int c = gt + q * 128;             // 2048 chunks
                const int row = c >> 5, ch = c & 31;
                const int dst_ch = (ch & ~3) | ((ch & 1) << 1) | ((ch >> 1) & 1);
                const int64_t gr = h0 + row;
                cp_async16(as_g + row * PAQ + dst_ch * 2, gr < rows ? Y + gr * ldy + ch * 2 : Y, gr < rows ? 16 : 0);
            }
        }
        cp_async_commit();
        // ---- C block of this warp: rows h0 + wm*32 + a*8 + lg, columns j0 + wn*32 + u*8 + 2 lk (+1)
        double2 nxt[4][4];
        auto load_c = [&](int t) {
            const int64_t j0 = (int64_t)t * NTQ + wn * 32 + 2 * lk;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int64_t h = h0 + wm * 32 + a * 8 + lg;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t j = j0 + u * 8;
                    nxt[a][u] = (h < rows && j < n) ? *reinterpret_cast<const double2*>(C + h * ldc + j) : make_double2(0., 0.);
                }
            }
        };
        if (live && t0 < t1) load_c(t0);
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        group_sync(g);                                  // the strip is staged
        for (int t = t0; t < t1; ++t, ++it) {
            const int s = it % QSTAGES; const uint32_t ph = (it / QSTAGES) & 1;
            double acc[4][4][2];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int u = 0; u < 4; ++u) { acc[a][u][0] = nxt[a][u].x; acc[a][u][1] = nxt[a][u].y; }
            if (live && t + 1 < t1) load_c(t + 1);
            mbar_wait(full + s, ph);
            if (live) {
                const double* ap = as_g + (wm * 32 + lg) * PAQ + lk;
                const double* bp = ring + (size_t)s * BTILE_DOUBLES + (wn * 4) * (KQ * 8);
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    double af[4], bf[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) af[a] = ap[a * 8 * PAQ + kk * 4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) bf[u] = bp[u * (KQ * 8) + koff[kk]];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int u = 0; u < 4; ++u) dmma884(acc[a][u][0], acc[a][u][1], af[a], bf[u]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
            if (live) {
                const int64_t j0 = (int64_t)t * NTQ + wn * 32 + 2 * lk;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int64_t h = h0 + wm * 32 + a * 8 + lg;
                    if (h >= rows) continue;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int64_t j = j0 + u * 8;
                        if (j < n) *reinterpret_cast<double2*>(C + h * ldc + j) = make_double2(acc[a][u][0], acc[a][u][1]);
                    }
                }
            }
        }
    }
}

}  // namespace

// returns PB_OK when the kernel ran, 1 when the shape / alignment is outside its domain (the caller falls back to the
// cp.async strip kernel), < 0 on error
int pbi_qr_update(pb_ctx* ctx, const double* Y, int64_t ldy, int64_t rows, int64_t K, const double* B, int64_t ldb, int64_t n,
                  double* C, int64_t ldc) {
    static const bool enabled = []{ const char* e = getenv("PB_QR_UPDATE_TMA"); return !e || atoi(e) != 0; }();
    if (!enabled || K != KQ || n < NTQ || (n % 8) || (ldb % 2) || (ldc % 2) || (ldy % 2)) return 1;
    if ((reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(Y)) % 16) return 1;
    if (rows < 128) return 1;
    CUtensorMap map;
    const uint64_t dims[3] = {8, (uint64_t)KQ, (uint64_t)(n / 8)};
    const uint64_t strides[2] = {(uint64_t)ldb * 8, 64};
    const uint32_t box[3] = {8, (uint32_t)KQ, NTQ / 8};
    if (pbi_tma_encode_3d(&map, B, dims, strides, box, 1) != 0) return 1;
    const int n_pairs = (int)((rows + 127) / 128);
    const int grid = std::min(n_pairs, ctx->sm_count);
    const int n_tiles = (int)((n + NTQ - 1) / NTQ);
    int csplit = 1; double best = 0.;
    // column splits only when there are too few strip pairs to balance the persistent grid by rows alone: a split costs a
    // second staging of the Y strip and a second pipeline fill per pair (measured: 250 000 x 4096 QR 354 -> 348 ms, C2
    // tsqr_full 23.2 -> 23.0 ms without them)
    for (int c = 1; c <= 4 && n_tiles / c >= 6 && (c == 1 || n_pairs < 2 * grid); ++c) {
        const int64_t units = (int64_t)n_pairs * c;
        const double eff = (double)units / (double)((units + grid - 1) / grid * grid);
        if (eff > best + 0.01) { best = eff; csplit = c; }
    }
    { static const int cs_env = getenv("PB_QR_UPDATE_CSPLIT") ? atoi(getenv("PB_QR_UPDATE_CSPLIT")) : 0; if (cs_env > 0) csplit = cs_env; }
    const size_t smem = sizeof(double) * ((size_t)QSTAGES * BTILE_DOUBLES + 2 * 64 * PAQ) + 2 * QSTAGES * sizeof(uint64_t) + 1024;
    PB_TRY(pbi_smem_optin(ctx, (const void*)qr_update_kernel, smem));
    qr_update_kernel<<<grid, QTHREADS, smem, ctx->stream>>>(map, Y, ldy, rows, n, C, ldc, n_pairs, csplit);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}
