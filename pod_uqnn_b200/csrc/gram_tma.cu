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
// (tile, k-iteration) space, every iteration weighted by what it costs the busiest SM sub-partition.
//   * no wave quantisation and no split heuristics: every SM gets the same weight;
//   * a tile is 16 x 16 "atoms" (one DMMA m8n8k4 accumulator each).  Off-diagonal interior tiles run the
//     plain 2 x 4 warp layout (64 x 32 per warp = 8 x 4 atoms, 64 atoms per sub-partition and iteration).
//     Diagonal tiles need only the atoms on or above the diagonal (136 of 256) and tiles of the last tile
//     column only the columns below n_s (13 of 16 atom columns at n_s = 1000): for those three kinds the
//     host cuts the needed 4 x 4-atom blocks into 8 chunks of nt <= 8 atom rows x 4 atom columns (nt A
//     fragments, 4 B fragments; the k loop is instantiated per nt, so the DMMA stream stays free of
//     predicates -- a per-atom mask measured 2 x slower) and pairs the chunks so that the four
//     sub-partitions (warps w and w + 4 share one) carry the same number of atoms: 40 / 64 for a diagonal
//     tile instead of the 48 / 64 of a quadrant skip, 52 / 64 for an edge tile (computed transposed, so
//     that the short side is the one split over the warp rows) instead of 64 / 64;
//   * the consumer warps are decoupled (no CTA barrier), so unequal chunks of the two warps of a
//     sub-partition overlap;
//   * the partition is static (host tables), so the partial tiles are summed in a fixed order.
enum { GK_FULL = 0, GK_EDGE = 1, GK_DIAG = 2, GK_DIAG_EDGE = 3 };
struct SegEntry { int ti, tj, k0, k1, slot, kind; };
struct WarpShape { int r0, c0, nt, pad; };      // nt atom rows from r0 x the 4 atom columns from c0
struct GramShapes { WarpShape w[4][8]; int transposed[4]; int weight[4]; };
struct GramPlan { int64_t ka = -1, rows = -1; int P = 0; int nslots = 0; std::vector<SegEntry> segs; std::vector<int> seg_begin,
                  tile_off, tile_src; int* d_tab = nullptr; size_t d_cap = 0; GramShapes shapes; long long* d_prof = nullptr; };

// k loop of one (tile, k range) segment for a warp that owns NT atom rows x 4 atom columns
template <int NT>
__device__ __forceinline__ void consume_segment(double (&acc)[8][4][2], const double* panels, uint64_t* full, uint64_t* empty,
                                                int a_off, int b_off, const int (&koff)[4], int k0, int k1, uint32_t& it, int lane) {
    for (int k = k0; k < k1; ++k, ++it) {
        const int s = it % TSTAGES; const uint32_t ph = (it / TSTAGES) & 1;
        mbar_wait(full + s, ph);
        if (NT > 0) {
            const double* st = panels + (size_t)s * 2 * PANEL_DOUBLES;
            const double* ap = st + a_off;
            const double* bp = st + b_off;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                double a[NT > 0 ? NT : 1], b[4];
#pragma unroll
                for (int t = 0; t < NT; ++t) a[t] = ap[t * (TBK * 8) + koff[kk]];
#pragma unroll
                for (int u = 0; u < 4; ++u) b[u] = bp[u * (TBK * 8) + koff[kk]];
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
gram_streamk_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ GramShapes shapes,
                    const SegEntry* __restrict__ segs, const int* __restrict__ seg_begin, double* __restrict__ ws,
                    long long* __restrict__ prof) {
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
    long long t_begin = 0;
    if (prof && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));

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
        // frame of this warp: nt atom rows from r0 (A fragments), 4 atom columns from c0 (B fragments)
        int r0 = wm * 8, c0 = wn * 4, nt = 8; bool tr = false;
        if (e.kind != GK_FULL) {
            const WarpShape sh = shapes.w[e.kind][warp];
            r0 = sh.r0; c0 = sh.c0; nt = sh.nt; tr = shapes.transposed[e.kind] != 0;
        }
        double acc[8][4][2];
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }
        const int a_off = ((tr && !diag) ? PANEL_DOUBLES : 0) + r0 * (TBK * 8);
        const int b_off = ((tr || diag) ? 0 : PANEL_DOUBLES) + c0 * (TBK * 8);
        // one straight-line k loop per row count: no predicates or branches between the DMMAs
        switch (nt) {
            case 8: consume_segment<8>(acc, panels, full, empty, a_off, b_off, koff, e.k0, e.k1, it, lane); break;
            case 7: consume_segment<7>(acc, panels, full, empty, a_off, b_off, koff, e.k0, e.k1, it, lane); break;
            case 6: consume_segment<6>(acc, panels, full, empty, a_off, b_off, koff, e.k0, e.k1, it, lane); break;
            case 5: consume_segment<5>(acc, panels, full, empty, a_off, b_off, koff, e.k0, e.k1, it, lane); break;
            case 4: consume_segment<4>(acc, panels, full, empty, a_off, b_off, koff, e.k0, e.k1, it, lane); break;
            case 3: consume_segment<3>(acc, panels, full, empty, a_off, b_off, koff, e.k0, e.k1, it, lane); break;
            case 2: consume_segment<2>(acc, panels, full, empty, a_off, b_off, koff, e.k0, e.k1, it, lane); break;
            case 1: consume_segment<1>(acc, panels, full, empty, a_off, b_off, koff, e.k0, e.k1, it, lane); break;
            default: consume_segment<0>(acc, panels, full, empty, a_off, b_off, koff, e.k0, e.k1, it, lane); break;
        }
        // partial tile -> workspace slot (dense 128 x 128; a transposed kind is stored transposed, the reduction knows)
        double* o = ws + (size_t)e.slot * (128 * 128);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (t < nt) {
                const int i = (r0 + t) * 8 + lg;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = (c0 + u) * 8 + 2 * lk;
                    *reinterpret_cast<double2*>(o + i * 128 + j) = make_double2(acc[t][u][0], acc[t][u][1]);
                }
            }
        }
    }
    if (prof && tid == 0) {                       // schedule calibration (pb_gram_profile): ns this CTA's first warp was busy
        long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        prof[blockIdx.x] = t_end - t_begin;
    }
}

// C[i][j] = sum over the slots of tile (ti, tj) in table order.  Everything below the diagonal is the mirror of
// the element above it (lower tiles do not exist, and diagonal tiles hold only their upper atoms), so C is exactly
// symmetric.  edge_tr: off-diagonal tiles of the last tile column were stored transposed.
__global__ void streamk_reduce_kernel(const double* __restrict__ ws, const int* __restrict__ tile_src_off,
                                      const int* __restrict__ tile_src, int ka, int tiles_n, int edge_tr,
                                      double* __restrict__ C, int64_t ldc) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)ka * ka) return;
    const int i = (int)(idx / ka), j = (int)(idx % ka);
    int si = i, sj = j;
    if (j < i) { si = j; sj = i; }
    const int ti = si / 128, tj = sj / 128;
    const int t = ti * tiles_n - ti * (ti - 1) / 2 + (tj - ti);
    int li = si % 128, lj = sj % 128;
    if (edge_tr && ti != tj && tj == tiles_n - 1) { const int x = li; li = lj; lj = x; }
    double s = 0.;
    for (int q = tile_src_off[t]; q < tile_src_off[t + 1]; ++q) s += ws[(size_t)tile_src[q] * (128 * 128) + li * 128 + lj];
    C[(int64_t)i * ldc + j] = s;
}

// ---- host: warp shapes of the three masked tile kinds
// need[r][c]: atom (r, c) of the frame orientation is wanted.  Per column block (4 atom columns) the wanted rows are
// cut into contiguous chunks of at most 8 rows; the number of chunks per block and the cuts minimise the atoms of
// the busiest sub-partition after pairing the chunks largest-with-smallest.
struct Chunk { int cb, ra, rb, atoms; };
bool plan_chunks(const bool need[16][16], std::vector<Chunk>& best, int& best_load) {
    int rlo[4], rhi[4], strip[4][16];
    for (int cb = 0; cb < 4; ++cb) {
        rlo[cb] = 16; rhi[cb] = 0;
        for (int r = 0; r < 16; ++r) {
            strip[cb][r] = 0;
            for (int u = 0; u < 4; ++u) if (need[r][cb * 4 + u]) strip[cb][r] = 4;      // a warp row is 4 atoms wide, wanted or not
            if (strip[cb][r]) { rlo[cb] = std::min(rlo[cb], r); rhi[cb] = std::max(rhi[cb], r + 1); }
        }
        if (rhi[cb] <= rlo[cb]) { rlo[cb] = rhi[cb] = 0; }
    }
    // split of rows [rlo, rhi) of block cb into n chunks (<= 8 rows each) minimising the largest chunk (DP over cut points)
    auto split = [&](int cb, int n, std::vector<Chunk>& out) -> bool {
        const int lo = rlo[cb], hi = rhi[cb], len = hi - lo;
        if (len == 0) return n == 0;
        if (n <= 0 || n * 8 < len || n > len) return false;
        int pre[17]; pre[0] = 0;
        for (int x = 0; x < len; ++x) pre[x + 1] = pre[x] + strip[cb][lo + x];
        const int INF = 1 << 30;
        std::vector<std::vector<int>> f(n + 1, std::vector<int>(len + 1, INF)), from(n + 1, std::vector<int>(len + 1, -1));
        f[0][0] = 0;
        for (int c = 1; c <= n; ++c)
            for (int x = 1; x <= len; ++x)
                for (int y = std::max(0, x - 8); y < x; ++y) {
                    if (f[c - 1][y] == INF) continue;
                    const int v = std::max(f[c - 1][y], pre[x] - pre[y]);
                    if (v < f[c][x]) { f[c][x] = v; from[c][x] = y; }
                }
        if (f[n][len] == INF) return false;
        std::vector<Chunk> tmp;
        for (int c = n, x = len; c >= 1; --c) { const int y = from[c][x]; tmp.push_back({cb, lo + y, lo + x, pre[x] - pre[y]}); x = y; }
        out.insert(out.end(), tmp.rbegin(), tmp.rend());
        return true;
    };
    best_load = 1 << 30; int best_single = 1 << 30; best.clear();
    int n[4];
    for (n[0] = 0; n[0] <= 8; ++n[0]) for (n[1] = 0; n[0] + n[1] <= 8; ++n[1]) for (n[2] = 0; n[0] + n[1] + n[2] <= 8; ++n[2])
        for (n[3] = 0; n[0] + n[1] + n[2] + n[3] <= 8; ++n[3]) {
            std::vector<Chunk> ch; bool ok = true;
            for (int cb = 0; cb < 4 && ok; ++cb) ok = split(cb, n[cb], ch);
            if (!ok) continue;
            while (ch.size() < 8) ch.push_back({0, 0, 0, 0});
            std::sort(ch.begin(), ch.end(), [](const Chunk& a, const Chunk& b) { return a.atoms > b.atoms; });
            int load = 0;
            for (int q = 0; q < 4; ++q) load = std::max(load, ch[q].atoms + ch[7 - q].atoms);
            if (load < best_load || (load == best_load && ch[0].atoms < best_single)) { best_load = load; best_single = ch[0].atoms; best = ch; }
        }
    return !best.empty();
}

// kinds for a matrix whose last tile column has cv valid atom columns (1 .. 16)
bool build_shapes(int cv, GramShapes& gs) {
    memset(&gs, 0, sizeof(gs));
    gs.weight[GK_FULL] = 640;
    for (int kind = 1; kind < 4; ++kind) {
        int best_load = 1 << 30, best_tr = 0; std::vector<Chunk> best;
        for (int tr = 0; tr < (kind == GK_EDGE ? 2 : 1); ++tr) {
            bool need[16][16];
            for (int r = 0; r < 16; ++r) for (int c = 0; c < 16; ++c) {
                const int ar = tr ? c : r, ac = tr ? r : c;      // atom of the tile (row of panel ti, column of panel tj)
                bool w = true;
                if (kind == GK_DIAG || kind == GK_DIAG_EDGE) w = w && (ar <= ac);
                if (kind == GK_EDGE || kind == GK_DIAG_EDGE) w = w && (ac < cv);
                need[r][c] = w;
            }
            std::vector<Chunk> ch; int load;
            if (!plan_chunks(need, ch, load)) continue;
            if (load < best_load) { best_load = load; best_tr = tr; best = ch; }
        }
        if (best.empty()) return false;
        gs.transposed[kind] = best_tr;
        // schedule weight of one 16-row iteration, in tenths of an atom: the atoms of the busiest sub-partition plus what
        // pb_gram_profile measured on top (per-stage hand-over and the unequal chunks of a sub-partition's two warps):
        // C2 on B200: full 2.13 us, edge 1.76, diagonal 1.41, diagonal + edge 1.41 per iteration = 640 / 528 / 423 / 424 (edge: 535 balanced best)
        gs.weight[kind] = 10 * std::max(best_load, 1) + (kind == GK_EDGE ? 15 : 23);
        if (kind == GK_DIAG_EDGE) gs.weight[kind] = std::max(gs.weight[kind], std::min(gs.weight[GK_DIAG], 10 * best_load + 54));
        // chunks are sorted by size: q and 7 - q share sub-partition q (warps q and q + 4)
        for (int q = 0; q < 8; ++q) {
            const Chunk& c = best[q < 4 ? q : 7 - (q - 4)];
            WarpShape& w = gs.w[kind][q];
            w.r0 = c.ra; w.c0 = c.cb * 4; w.nt = c.rb - c.ra;
        }
    }
    return true;
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

extern "C" int pb_gram_tile_shapes(int valid_atom_cols, int out[3][8][3], int transposed[3], int weight[3]) {
    if (valid_atom_cols < 1 || valid_atom_cols > 16 || !out || !transposed || !weight) return PB_ERR_ARG;
    GramShapes gs;
    if (!build_shapes(valid_atom_cols, gs)) return PB_ERR_ARG;
    for (int k = 1; k < 4; ++k) {
        for (int w = 0; w < 8; ++w) { out[k - 1][w][0] = gs.w[k][w].r0; out[k - 1][w][1] = gs.w[k][w].c0; out[k - 1][w][2] = gs.w[k][w].nt; }
        transposed[k - 1] = gs.transposed[k]; weight[k - 1] = gs.weight[k];
    }
    return PB_OK;
}

// Schedule calibration: after pb_gram_profile(ctx, 1, ...) the stream-K kernel records the busy time of every CTA; a later
// call returns, per CTA, that time and the k iterations it ran of each tile kind (full, edge, diagonal, diagonal + edge).
extern "C" int pb_gram_profile(pb_ctx* ctx, int enable, int capacity, double* busy_us, int* iters /* [capacity][4] */, int* n_cta) {
    PB_ENTER(ctx);
    if (!ctx->gram_plan) ctx->gram_plan = new GramPlan();
    GramPlan& plan = *static_cast<GramPlan*>(ctx->gram_plan);
    if (enable && !plan.d_prof) { PB_CUDA(ctx, cudaMalloc(&plan.d_prof, sizeof(long long) * 1024)); PB_CUDA(ctx, cudaMemset(plan.d_prof, 0, sizeof(long long) * 1024)); }
    if (!enable && plan.d_prof) { PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(plan.d_prof); plan.d_prof = nullptr; }
    if (n_cta) *n_cta = plan.ka > 0 ? plan.P : 0;
    if (busy_us && iters && plan.d_prof && plan.ka > 0) {
        if (capacity < plan.P || plan.P > 1024) return PB_ERR_ARG;
        std::vector<long long> t(plan.P);
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PB_CUDA(ctx, cudaMemcpy(t.data(), plan.d_prof, sizeof(long long) * plan.P, cudaMemcpyDeviceToHost));
        for (int c = 0; c < plan.P; ++c) {
            busy_us[c] = (double)t[c] * 1e-3;
            for (int k = 0; k < 4; ++k) iters[c * 4 + k] = 0;
            for (int sg = plan.seg_begin[c]; sg < plan.seg_begin[c + 1]; ++sg) iters[c * 4 + plan.segs[sg].kind] += plan.segs[sg].k1 - plan.segs[sg].k0;
        }
    }
    return PB_OK;
}

void pbi_gram_plan_free(pb_ctx* ctx) {
    if (!ctx->gram_plan) return;
    GramPlan* p = static_cast<GramPlan*>(ctx->gram_plan);
    if (p->d_tab) cudaFree(p->d_tab);
    if (p->d_prof) cudaFree(p->d_prof);
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


// stream-K launcher: returns PB_OK when it produced C (both triangles, exactly symmetric), 1 to fall back
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
        plan.ka = -1; plan.rows = rows; plan.P = P; plan.segs.clear(); plan.seg_begin.assign(P + 1, 0);
        const int cv = (int)((ka - (int64_t)(tiles_n - 1) * 128) / 8);          // valid atom columns of the last tile column
        if (!build_shapes(cv, plan.shapes)) return 1;
        if (const char* e = getenv("PB_GRAM_WEIGHTS")) {          // experiment knob: "edge,diag,diag_edge" schedule weights (a full tile is 640)
            int w1, w2, w3;
            if (sscanf(e, "%d,%d,%d", &w1, &w2, &w3) == 3) { plan.shapes.weight[GK_EDGE] = w1; plan.shapes.weight[GK_DIAG] = w2; plan.shapes.weight[GK_DIAG_EDGE] = w3; }
        }
        const GramShapes& gs = plan.shapes;
        auto kind_of = [&](int ti, int tj) {
            const bool edge = (cv < 16 && tj == tiles_n - 1);
            return ti == tj ? (edge ? GK_DIAG_EDGE : GK_DIAG) : (edge ? GK_EDGE : GK_FULL);
        };
        auto wt = [&](int ti, int tj) { return (int64_t)std::max(gs.weight[kind_of(ti, tj)], 160); };   // floor: a stage is still 16-32 KB of TMA traffic
        const double WF = (double)gs.weight[GK_FULL];
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
        double Wtot = 0.;
        for (auto& t : off) Wtot += (double)nk * wt(t.first, t.second);
        for (auto& t : dia) Wtot += (double)nk * wt(t.first, t.second);
        const double D = Wtot / P;
        if (n_off > 0 && WF * nk >= D) {                       // pieces of tiles, one round
            const int64_t L = std::max<int64_t>(1, (int64_t)(D / WF));
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
        auto wof = [&](const Item& it) { return (it.k1 - it.k0) * wt(it.ti, it.tj); };
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
            x = lo; k = (b - cum[lo]) / wt(tail[lo].ti, tail[lo].tj);
        };
        const int Tn = tiles_n;
        auto tile_id = [&](int ti, int tj) { return ti * Tn - ti * (ti - 1) / 2 + (tj - ti); };
        std::vector<std::vector<int>> src(T);
        int slot = 0;
        double acc_need = 0.;
        for (int sI = 0; sI < P; ++sI) {
            plan.seg_begin[sI] = (int)plan.segs.size();
            for (auto& it : own[sI]) {
                plan.segs.push_back({it.ti, it.tj, (int)it.k0, (int)it.k1, slot, kind_of(it.ti, it.tj)});
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
                    plan.segs.push_back({tail[x].ti, tail[x].tj, (int)ks, (int)ke, slot, kind_of(tail[x].ti, tail[x].tj)});
                    src[tile_id(tail[x].ti, tail[x].tj)].push_back(slot); ++slot;
                }
            }
        }
        plan.seg_begin[P] = (int)plan.segs.size();
        plan.nslots = slot;
        plan.tile_off.assign(T + 1, 0); plan.tile_src.clear();
        for (int tt = 0; tt < T; ++tt) { plan.tile_off[tt] = (int)plan.tile_src.size(); for (int v : src[tt]) plan.tile_src.push_back(v); }
        plan.tile_off[T] = (int)plan.tile_src.size();
        // device tables: [segs (6 ints each)] [seg_begin (P+1)] [tile_off (T+1)] [tile_src]
        const size_t n_int = plan.segs.size() * 6 + (P + 1) + (T + 1) + plan.tile_src.size();
        if (plan.d_cap < n_int) { if (plan.d_tab) cudaFree(plan.d_tab); PB_CUDA(ctx, cudaMalloc(&plan.d_tab, n_int * sizeof(int))); plan.d_cap = n_int; }
        std::vector<int> host(n_int); size_t o = 0;
        memcpy(host.data(), plan.segs.data(), plan.segs.size() * 6 * sizeof(int)); o += plan.segs.size() * 6;
        memcpy(host.data() + o, plan.seg_begin.data(), (P + 1) * sizeof(int)); o += P + 1;
        memcpy(host.data() + o, plan.tile_off.data(), (T + 1) * sizeof(int)); o += T + 1;
        memcpy(host.data() + o, plan.tile_src.data(), plan.tile_src.size() * sizeof(int));
        PB_CUDA(ctx, cudaMemcpyAsync(plan.d_tab, host.data(), n_int * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        plan.ka = ka;
    }
    const SegEntry* d_segs = reinterpret_cast<const SegEntry*>(plan.d_tab);
    const int* d_seg_begin = plan.d_tab + plan.segs.size() * 6;
    const int* d_tile_off = d_seg_begin + (P + 1);
    const int* d_tile_src = d_tile_off + (T + 1);
    PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)plan.nslots * 128 * 128));
    const size_t smem = (size_t)TSTAGES * STAGE_BYTES + 2 * TSTAGES * sizeof(uint64_t) + 1024;
    PB_TRY(pbi_smem_optin(ctx, (const void*)gram_streamk_kernel, (size_t)(smem)));
    if (time_it) cudaEventRecord(ctx->evk0, ctx->stream);
    gram_streamk_kernel<<<P, TMA_THREADS, smem, ctx->stream>>>(map, plan.shapes, d_segs, d_seg_begin, ctx->ws, plan.d_prof);
    PB_LAUNCH_CHECK(ctx);
    if (time_it) { cudaEventRecord(ctx->evk1, ctx->stream); ctx->tn_timing_pending = true; }
    const int64_t n = ka * ka;
    streamk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->ws, d_tile_off, d_tile_src, (int)ka, tiles_n,
                                                                                         (ka % 128) ? plan.shapes.transposed[GK_EDGE] : 0, C, ldc);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}
