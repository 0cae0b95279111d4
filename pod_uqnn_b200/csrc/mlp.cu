// Deep-ensemble surrogate forward: M mean/variance MLPs evaluated on a batch of parameter
// points, fused with the Gaussian-mixture moments of PodnnModel.predict_v.
//
// Replaces varneuralnetwork.py:42-62 (Dense+ReLU stack, linear 2*n_L head, softplus scale),
// :75-86 (input normalisation), :154-160 (mean / variance) and podnnmodel.py:314-328
// (mean over members of mu and of sigma^2 + mu^2) -- serial over members and layers in the
// reference, with (N, n_L, M) intermediates.
//
// One persistent CTA owns a tile of 128 points and walks all members and layers for it:
//   * activations stay in shared memory ([128][132] doubles, rewritten in place per layer);
//   * the weights are repacked once per call into the order the kernel consumes them
//     (16-row K chunks padded to a 132-double pitch, each layer followed by its bias row), so a
//     PRODUCER warp streams them with plain 1-D bulk async copies (cp.async.bulk, SASS UBLKCP)
//     into a 4-slot ring guarded by full / empty mbarriers and runs ahead across layer and member
//     boundaries -- the consumers never wait for a pipeline refill at a layer start;
//   * 8 CONSUMER warps run the layer product on DMMA (warp tile 64 x 32, accumulators in
//     registers), apply bias + ReLU in the epilogue and meet only at two named barriers per layer;
//   * the head epilogue stores mu and sigma^2 of the member to a per-CTA scratch (L2 resident);
//     after the last member the tile is reduced to (mean, sqrt(var)).  No (N, n_L, M)
//     intermediate reaches HBM as such, and nothing is read-modify-written across members.
#include "common.cuh"
#include "tile_ops.cuh"
#include <cstdlib>
#include <algorithm>

namespace {
using namespace pbt;

constexpr int PACT = 132;          // activation pitch (== 4 mod 16: conflict-free A fragments)
constexpr int PW = 132;            // packed weight pitch
constexpr int RING = 7;             // weight ring slots (96-point tile: 101 KB activations + 118 KB ring)
constexpr int CHUNK_DOUBLES = BK * PW;             // 2112 doubles = 16 896 B per K chunk
constexpr int BIAS_DOUBLES = PW;                   // 132 doubles = 1 056 B
constexpr int MAX_LAYERS = 8;
constexpr int MAX_WIDTH = 128;     // hidden widths (inputs of every layer) must fit the tile
constexpr int MLP_GROUPS = 3;      // consumer groups of 4 warps, 32 points each (tile = 96 points); see ensemble_kernel (PB_MLP_GROUPS=1|2|3)

struct MlpArgs {
    int n_members, n_layers, n_L, norm_kind;
    int dims[MAX_LAYERS + 1];
    int64_t member_stride;          // doubles per member in the packed stream
    const double* packed; const double* norm_a; const double* norm_b;
    double soft0;
    const double* X; int64_t N;
    double* v_mean; double* v_sig;
    double* scratch;                // per CTA: [member][MP][2*n_L] raw head outputs t = h W + b
    int64_t n_tiles;
    unsigned skew_ns;
};

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
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the consumer groups (4 warps = 32 points of the tile each) only synchronise internally, so one group's
// epilogue / mixture pass overlaps the other groups' DMMA work
__device__ __forceinline__ void group_sync(int bar_id) { asm volatile("bar.sync %0, 128;\n" :: "r"(bar_id) : "memory"); }

// max(v, 0) on the integer pipe (sign bit of the high word): -0 and negative NaNs become +0
__device__ __forceinline__ double relu_bits(double v) { return __double2hiint(v) < 0 ? 0. : v; }

__device__ __forceinline__ double softplus(double x) {
    // np.logaddexp(0, x) == tf.math.softplus within 1 ulp.
    // |x| <= 1/4 (the usual range: soft_0 = 0.01 times an O(1) pre-activation): ln(1 + e^x) =
    // ln 2 + x/2 + ln cosh(x/2), with ln cosh(y) = y^2/2 - y^4/12 + y^6/45 - 17 y^8/2520 + 31 y^10/14175
    // - 691 y^12/935550 + 10922 y^14/42567525; truncation < 2e-16 relative (checked against long double),
    // 9 FMAs instead of the ~100 fp64 instructions of exp + log1p, which share the pipe with the DMMAs.
    if (fabs(x) <= 0.25) {
        const double y2 = 0.25 * x * x;
        double p = 10922. / 42567525.;
        p = fma(p, y2, -691. / 935550.);
        p = fma(p, y2, 31. / 14175.);
        p = fma(p, y2, -17. / 2520.);
        p = fma(p, y2, 1. / 45.);
        p = fma(p, y2, -1. / 12.);
        p = fma(p, y2, 0.5);
        return fma(p, y2, fma(0.5, x, 0.6931471805599453));
    }
    return x > 0. ? x + log1p(exp(-x)) : log1p(exp(x));
}

// Keras layout (per member: W0 (in x out) row-major, b0, W1, b1, ...) -> packed stream
// (per member, per layer, per 128-wide output chunk: bias [132], then ceil(in/16) blocks [16][132])
__global__ void pack_weights_kernel(const double* __restrict__ w, double* __restrict__ packed, int n_members,
                                    int n_layers, MlpArgs a, int64_t src_member_stride) {
    const int64_t total = (int64_t)n_members * a.member_stride;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int mem = (int)(idx / a.member_stride);
        int64_t r = idx % a.member_stride;
        int64_t src_off = 0;
        double v = 0.;
        for (int l = 0; l < n_layers; ++l) {
            const int in = a.dims[l], out = a.dims[l + 1];
            const int nk = (in + BK - 1) / BK, nc = (out + 127) / 128;
            const int64_t per_chunk = (int64_t)nk * CHUNK_DOUBLES + BIAS_DOUBLES;
            const int64_t layer_sz = per_chunk * nc;
            if (r < layer_sz) {
                const int c = (int)(r / per_chunk); const int64_t q = r % per_chunk;
                const double* W = w + (int64_t)mem * src_member_stride + src_off;
                if (q >= BIAS_DOUBLES) {
                    const int64_t qq = q - BIAS_DOUBLES;
                    const int kc = (int)(qq / CHUNK_DOUBLES); const int e = (int)(qq % CHUNK_DOUBLES);
                    const int row = kc * BK + e / PW, col = c * 128 + e % PW;
                    if (row < in && (e % PW) < 128 && col < out) v = W[(int64_t)row * out + col];
                } else {
                    const int e = (int)q; const int col = c * 128 + e;
                    if (e < 128 && col < out) v = W[(int64_t)in * out + col];
                }
                break;
            }
            r -= layer_sz;
            src_off += (int64_t)in * out + out;
        }
        packed[idx] = v;
    }
}

// NG consumer groups of 4 warps; a group owns 32 points of the tile (warp tile 32 points x 32 outputs).  Every SM
// sub-partition hosts one warp of each group, so while one group is in a layer epilogue / the mixture pass / the input
// stage the other groups' warps keep its DMMA pipe fed.  (Two groups of 64 points measured 72 % tensor-pipe activity:
// a group spends ~30 % of its time outside the products, and ONE warp per sub-partition reaches only ~60 % of the pipe.)
// Block = NG * 4 consumer warps + ONE producer warp (13 warps at NG = 3, 128 registers per thread).  A 16-warp layout
// (producer warpgroup + setmaxnreg 40 / 152, i.e. 512 threads x 128 registers = the whole register file) was 3 % faster
// but produced sporadic wrong tiles -- one consumer warp's 32 output columns of one member, ~10 of 3 000 tiles, only with
// >= 5 ring slots, with or without setmaxnreg; NG = 1, 2 and this 13-warp layout passed 12 launches x 1e6 points against
// the oracle (tools/stress_ens.py, tests/test_gpu_scale.py::test_ensemble_predict_repeated_launches).  Root cause not found.
template <int NG>
__global__ void __launch_bounds__((NG * 4 + 1) * 32, 1) ensemble_kernel(MlpArgs a) {
    constexpr int MPT = 32 * NG;                          // points per tile
    constexpr int NCW = NG * 4;                           // consumer warps
    extern __shared__ __align__(16) double smem[];
    double* act = smem;                                   // [MPT][PACT]
    double* ring = smem + MPT * PACT;                     // [RING][CHUNK_DOUBLES]
    double* xs = ring + RING * CHUNK_DOUBLES;             // [MPT][4]: normalised inputs of the tile when n_d <= 4
    uint64_t* full = reinterpret_cast<uint64_t*>(xs + MPT * 4);
    uint64_t* empty = full + RING;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_L = a.n_L;
    if (tid == 0) {
        for (int s = 0; s < RING; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    if (warp >= NCW) {
        // ------------------------------------------------ producer: streams the packed weights
        if (warp == NCW && lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x)
                for (int mem = 0; mem < a.n_members; ++mem) {
                    const double* src = a.packed + (int64_t)mem * a.member_stride;
                    for (int l = 0; l < a.n_layers; ++l) {
                        const int nk = (a.dims[l] + BK - 1) / BK, nc = (a.dims[l + 1] + 127) / 128;
                        for (int c = 0; c < nc; ++c)
                            for (int kc = -1; kc < nk; ++kc, ++it) {       // kc == -1: the bias block (it seeds the accumulators)
                                const int s = it % RING; const uint32_t ph = (it / RING) & 1;
                                const uint32_t bytes = (kc >= 0 ? CHUNK_DOUBLES : BIAS_DOUBLES) * 8;
                                mbar_wait(empty + s, ph ^ 1);
                                mbar_expect_tx(full + s, bytes);
                                bulk_load(ring + (size_t)s * CHUNK_DOUBLES, src, bytes, full + s);
                                src += bytes / 8;
                            }
                    }
                }
        }
        return;
    }

    // ---------------------------------------------------- consumers
    const int g = warp >> 2, wn = warp & 3;               // group, 32-column slice of the layer output
    const int gt = tid & 127;                             // thread inside the group
    const int bid = 1 + g;                                // named barrier of the group
    const int lk = lane & 3, lg = lane >> 2;
    const int n_d = a.dims[0];
    const bool small_in = (n_d <= 4);                     // first layer on plain FMAs from xs (no 16-column padding)
    const int out_head = 2 * n_L;
    double* scr = a.scratch + (size_t)blockIdx.x * a.n_members * MPT * out_head;
    uint32_t it = 0;
    // Skew the groups against each other so that their epilogues do not coincide (all groups consume the same weight
    // ring, which bounds the skew to its depth).
    if (g > 0 && a.skew_ns > 0) __nanosleep(a.skew_ns * g);
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * MPT;
        if (small_in) {
            // ---- inputs of the tile, normalised once for all members (varneuralnetwork.py:75-86)
            if (gt < 32 * 4) {
                const int p = g * 32 + (gt >> 2), c = gt & 3;
                double v = 0.;
                if (c < n_d && p0 + p < a.N) {
                    v = a.X[(p0 + p) * n_d + c];
                    if (a.norm_kind == PB_NORM_MEANSTD) v = (v - a.norm_a[c]) / a.norm_b[c];
                    else if (a.norm_kind == PB_NORM_CENTER) v = (v - a.norm_a[c]) - 0.5 * (a.norm_b[c] - a.norm_a[c]);
                }
                xs[p * 4 + c] = v;
            }
            group_sync(bid);
        }
        for (int mem = 0; mem < a.n_members; ++mem) {
            if (!small_in) {
                // ---- input tile, normalised, zero padded to a multiple of 16 columns
                const int kin0 = (n_d + BK - 1) / BK * BK;
                for (int e = gt; e < 32 * kin0; e += 128) {
                    const int p = g * 32 + e / kin0, c = e % kin0;
                    double v = 0.;
                    if (c < n_d && p0 + p < a.N) {
                        v = a.X[(p0 + p) * n_d + c];
                        if (a.norm_kind == PB_NORM_MEANSTD) v = (v - a.norm_a[c]) / a.norm_b[c];
                        else if (a.norm_kind == PB_NORM_CENTER) v = (v - a.norm_a[c]) - 0.5 * (a.norm_b[c] - a.norm_a[c]);
                    }
                    act[p * PACT + c] = v;
                }
                group_sync(bid);
            }
            for (int l = 0; l < a.n_layers; ++l) {
                const int in = a.dims[l], out = a.dims[l + 1];
                const bool head = (l == a.n_layers - 1);
                const int nk = (in + BK - 1) / BK, nc = (out + 127) / 128;
                for (int c = 0; c < nc; ++c) {                 // > 1 chunk only for a wide head
                    // the bias block of this (layer, chunk) comes first and seeds the accumulators: the epilogue then has no
                    // fp64-pipe instruction at all (a DADD queues behind the other groups' DMMAs for ~40 cycles each)
                    double acc[4][4][2];
                    {
                        const int sb = it % RING; const uint32_t phb = (it / RING) & 1;
                        ++it;
                        mbar_wait(full + sb, phb);
                        const double* bias = ring + (size_t)sb * CHUNK_DOUBLES + wn * 32 + 2 * lk;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double2 b2 = *reinterpret_cast<const double2*>(bias + u * 8);
#pragma unroll
                            for (int t = 0; t < 4; ++t) { acc[t][u][0] = b2.x; acc[t][u][1] = b2.y; }
                        }
                        // the slot may only be released once the loads have RETURNED: nothing else consumes them before the arrive
                        asm volatile("" :: "d"(acc[0][0][0]), "d"(acc[0][1][0]), "d"(acc[0][2][0]), "d"(acc[0][3][1]) : "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty + sb);
                    }
                    if (l == 0 && small_in) {
                        // x W0 for 1..4 inputs: n_d FMAs per output from the one weight chunk (rows >= n_d are zero padding)
                        const int s = it % RING; const uint32_t ph = (it / RING) & 1;
                        ++it;
                        mbar_wait(full + s, ph);
                        const double* w0 = ring + (size_t)s * CHUNK_DOUBLES + wn * 32 + 2 * lk;
                        for (int ci = 0; ci < n_d; ++ci) {
                            double xv[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) xv[t] = xs[(g * 32 + t * 8 + lg) * 4 + ci];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const double2 w = *reinterpret_cast<const double2*>(w0 + ci * PW + u * 8);
#pragma unroll
                                for (int t = 0; t < 4; ++t) { acc[t][u][0] = fma(xv[t], w.x, acc[t][u][0]); acc[t][u][1] = fma(xv[t], w.y, acc[t][u][1]); }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty + s);
                    } else {
#pragma unroll 1
                        for (int kt = 0; kt < nk; ++kt, ++it) {
                            const int s = it % RING; const uint32_t ph = (it / RING) & 1;
                            mbar_wait(full + s, ph);
                            const double* as = act + (g * 32 + lg) * PACT + kt * BK + lk;
                            const double* bs = ring + (size_t)s * CHUNK_DOUBLES + wn * 32 + lg;
#pragma unroll
                            for (int kk = 0; kk < BK / 4; ++kk) {
                                double af[4], bf[4];
#pragma unroll
                                for (int t = 0; t < 4; ++t) af[t] = as[t * 8 * PACT + kk * 4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) bf[u] = bs[(kk * 4 + lk) * PW + u * 8];
#pragma unroll
                                for (int t = 0; t < 4; ++t)
#pragma unroll
                                    for (int u = 0; u < 4; ++u) dmma884(acc[t][u][0], acc[t][u][1], af[t], bf[u]);
                            }
                            __syncwarp();
                            if (lane == 0) mbar_arrive(empty + s);
                        }
                    }
                    if (!head) {
                        if (!(l == 0 && small_in)) group_sync(bid);   // every warp of the group has finished reading act for this layer
                        // next activations: relu(acc + b), columns [out, ceil16(out)) zeroed
                        const int outp = (out + BK - 1) / BK * BK;
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int p = g * 32 + t * 8 + lg;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int n = wn * 32 + u * 8 + 2 * lk;
                                double2 v;
                                v.x = n < out ? relu_bits(acc[t][u][0]) : 0.;
                                v.y = n + 1 < out ? relu_bits(acc[t][u][1]) : 0.;
                                if (n < outp) *reinterpret_cast<double2*>(act + p * PACT + n) = v;
                            }
                        }
                        group_sync(bid);            // activations of the next layer are complete
                    } else {
                        // head: raw outputs t = h W + b go to the per-CTA scratch (L2); the Normal's
                        // parameters are formed in the mixture pass below
                        double* sm = scr + (size_t)mem * MPT * out_head;
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int p = g * 32 + t * 8 + lg;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int nl = wn * 32 + u * 8 + 2 * lk;          // column inside the chunk (even)
                                const int n = c * 128 + nl;
                                if (n + 1 < out)
                                    *reinterpret_cast<double2*>(sm + (size_t)p * out_head + n) = make_double2(acc[t][u][0], acc[t][u][1]);
                                else if (n < out) sm[(size_t)p * out_head + n] = acc[t][u][0];
                            }
                        }
                    }
                }
            }
            group_sync(bid);        // act may be overwritten by the next member's input tile / first layer
        }
        // ---- mixture moments (podnnmodel.py:322-326): mean_i mu_i, sqrt(mean_i(var_i + mu_i^2) - mean^2)
        __threadfence();          // the head outputs are re-read through L2 (ld.global.cg) by other threads of the group
        group_sync(bid);
        const double invM = 1. / (double)a.n_members;
        const int64_t pts = min((int64_t)32, a.N - p0 - g * 32);      // points of this group
        // 4 members x 2 elements per pass: 16 independent loads, then 8 independent softplus chains (the fp64
        // exp / log1p sequences are latency bound one at a time; measured 30 % of the kernel before)
        const int64_t n_el = pts * n_L;
#pragma unroll 1
        for (int64_t e0 = gt; e0 < n_el; e0 += 256) {
            int64_t el[2] = {e0, e0 + 128};
            const bool ok1 = el[1] < n_el;
            if (!ok1) el[1] = e0;
            double s1[2] = {0., 0.}, s2[2] = {0., 0.}, var1[2] = {0., 0.};
            const double* base[2]; int jj[2];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const int64_t e = (int64_t)g * 32 * n_L + el[x];
                const int p = (int)(e / n_L); jj[x] = (int)(e % n_L);
                base[x] = scr + (size_t)p * out_head;
            }
#pragma unroll 1
            for (int m0 = 0; m0 < a.n_members; m0 += 4) {
                double mu4[2][4], rw4[2][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int mem = min(m0 + k, a.n_members - 1);
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        const double* tm = base[x] + (size_t)mem * MPT * out_head;
                        // L2 loads (ld.global.cg): the scratch is rewritten every tile by this CTA's own STGs, and a line
                        // still resident in L1 from the previous tile's pass was observed to be served stale (sporadic 5 %
                        // errors in a few of 3 000 tiles) -- the data is read once, L1 has nothing to offer here
                        mu4[x][k] = __ldcg(tm + jj[x]);                     // loc = t[:, :n_L]   (:57)
                        rw4[x][k] = __ldcg(tm + n_L + jj[x]);
                    }
                }
                double sc4[2][4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int x = 0; x < 2; ++x) sc4[x][k] = softplus(a.soft0 * rw4[x][k]) + 1e-6;   // scale (:58)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (m0 + k < a.n_members) {
#pragma unroll
                        for (int x = 0; x < 2; ++x) {
                            var1[x] = sc4[x][k] * sc4[x][k];
                            s1[x] += mu4[x][k]; s2[x] += var1[x] + mu4[x][k] * mu4[x][k];
                        }
                    }
                }
            }
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                if (x == 1 && !ok1) break;
                const double mean = s1[x] * invM;
                // one member: its own variance, exactly (VarNeuralNetwork.predict, varneuralnetwork.py:159)
                const double var = a.n_members == 1 ? var1[x] : s2[x] * invM - mean * mean;
                const int64_t o = (p0 + (int64_t)g * 32) * n_L + el[x];
                a.v_mean[o] = mean;
                a.v_sig[o] = sqrt(var);
            }
        }
        group_sync(bid);            // scratch is reused by the next tile
    }
}

// ---- Monte-Carlo version of PodnnModel.predict (podnnmodel.py:366-379) -----------------------------
// Philox4x32-10 (counter based: draw, point, mode -> 128 random bits), Box-Muller to N(0, 1).
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// vT[j][p] = v_mean[p][j] + v_sig[p][j] * z(draw, p, j)      (n_L x ldb, the B operand of V v^T)
__global__ void sample_v_kernel(const double* __restrict__ v_mean, const double* __restrict__ v_sig, int64_t N, int n_L,
                                uint64_t seed, uint32_t draw, double* __restrict__ vT, int64_t ldb) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // pair index: two normals per Philox call
    const int64_t total = N * n_L;
    if (2 * idx >= total) return;
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), draw, 0x5eedu};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    // two uniforms in (0, 1) with 53 / 32 bits, Box-Muller
    const double u1 = ((double)(((uint64_t)c[0] << 21) ^ (c[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)c[2] + 0.5) * (1.0 / 4294967296.0);
    const double rad = sqrt(-2. * log(u1));
    double sn, cs; sincospi(2. * u2, &sn, &cs);
    const double z[2] = {rad * cs, rad * sn};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int64_t e = 2 * idx + q;
        if (e < total) {
            const int64_t p = e / n_L; const int j = (int)(e % n_L);
            vT[(int64_t)j * ldb + p] = v_mean[e] + v_sig[e] * z[q];
        }
    }
}
// U_pred = sum / S ; U_sig = sqrt((S sumsq - sum^2) / (S (S - 1)))     (podnnmodel.py:376-378)
__global__ void mc_finalize_kernel(double* __restrict__ sum, double* __restrict__ sq, int64_t rows, int64_t cols, int64_t ld,
                                   double S) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t o = (i / cols) * ld + i % cols;
    const double a = sum[o], b = sq[o];
    sum[o] = a / S;
    sq[o] = sqrt(fmax(S * b - a * a, 0.) / (S * (S - 1.)));
}

// Gaussian mixture of the members at the U level (podnnmodel.py:344-362): running sum of the member means and of
// sig_i^2 + U_i^2, then U_pred = mean, U_sig = sqrt(mean(sig^2 + U^2) - U_pred^2) (+ pod_sig per row)
__global__ void u_mixture_accumulate_kernel(const double* __restrict__ Um, const double* __restrict__ Us, int64_t ldi,
                                          double* __restrict__ sum, double* __restrict__ m2, int64_t lda,
                                          int64_t rows, int64_t cols, int first) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols, c = i % cols;
    const double u = Um[r * ldi + c], sg = Us[r * ldi + c];
    const double q = sg * sg + u * u;
    const int64_t o = r * lda + c;
    sum[o] = first ? u : sum[o] + u;
    m2[o] = first ? q : m2[o] + q;
}
__global__ void u_mixture_finalize_kernel(double* __restrict__ sum, double* __restrict__ m2, int64_t ld, int64_t rows,
                                        int64_t cols, double n_members, const double* __restrict__ pod_sig) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols, o = r * ld + i % cols;
    const double mean = sum[o] / n_members;
    double sg = sqrt(m2[o] / n_members - mean * mean);      // NaN for a negative rounding residue, as np.sqrt
    if (pod_sig) sg += pod_sig[r];
    sum[o] = mean; m2[o] = sg;
}

// closed-form U-level moments: U_mean = V v_mean^T ; U_sig = sqrt((V*V) (v_sig^2)^T)
__global__ void square_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] * in[i];
}
__global__ void square_strided_kernel(const double* __restrict__ in, int64_t ld, double* __restrict__ out,
                                      int64_t rows, int cols) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int64_t r = i / cols; int c = (int)(i % cols);
    double v = in[r * ld + c]; out[i] = v * v;
}
__global__ void sqrt_inplace_kernel(double* __restrict__ x, int64_t rows, int64_t cols, int64_t ld) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int64_t r = i / cols, c = i % cols;
    x[r * ld + c] = sqrt(fmax(x[r * ld + c], 0.));
}

// ---- layer-by-layer path for nets with a layer wider than the fused kernel's 128-column activation tile (e.g. the
//      reference's 1dt_shallowwater h_layers = [256, 256, 256]): Dense products on the generic DMMA NN kernel,
//      bias / ReLU and the mixture moments as elementwise passes.  Same formulas as ensemble_kernel.
__global__ void normalize_rows_kernel(const double* __restrict__ X, int64_t n, int d, int kind, const double* __restrict__ na,
                                      const double* __restrict__ nb, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * d) return;
    const int c = (int)(i % d);
    double v = X[i];
    if (kind == PB_NORM_MEANSTD) v = (v - na[c]) / nb[c];
    else if (kind == PB_NORM_CENTER) v = (v - na[c]) - 0.5 * (nb[c] - na[c]);
    out[i] = v;
}
__global__ void bias_act_kernel(double* __restrict__ H, int64_t n, int cols, const double* __restrict__ bias, int relu) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * cols) return;
    const double v = H[i] + bias[i % cols];
    H[i] = relu ? fmax(v, 0.) : v;
}
// T (n x 2 n_L): loc = T[:, :n_L], scale = softplus(soft0 T[:, n_L:]) + 1e-6 (varneuralnetwork.py:57-58)
__global__ void mixture_accum_kernel(const double* __restrict__ T, int64_t n, int n_L, double soft0, int first,
                                     double* __restrict__ s1, double* __restrict__ s2, double* __restrict__ var1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n_L) return;
    const int64_t p = i / n_L; const int j = (int)(i % n_L);
    const double mu = T[p * 2 * n_L + j];
    const double sc = softplus(soft0 * T[p * 2 * n_L + n_L + j]) + 1e-6;
    const double var = sc * sc;
    s1[i] = (first ? 0. : s1[i]) + mu;
    s2[i] = (first ? 0. : s2[i]) + (var + mu * mu);
    var1[i] = var;
}
__global__ void mixture_final_kernel(double* __restrict__ s1, double* __restrict__ s2, const double* __restrict__ var1,
                                     int64_t n, int M) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mean = s1[i] / (double)M;
    const double var = M == 1 ? var1[i] : s2[i] / (double)M - mean * mean;      // (podnnmodel.py:322-326)
    s1[i] = mean; s2[i] = sqrt(var);
}

int ensemble_predict_layered(pb_ctx* ctx, int M, int n_layers, const int* dims, const double* w, int norm_kind,
                             const double* na, const double* nb, double soft0, const double* X, int64_t N,
                             double* v_mean, double* v_sig) {
    const int n_L = dims[n_layers] / 2;
    int maxw = 0; int64_t per_member = 0;
    for (int l = 0; l <= n_layers; ++l) maxw = std::max(maxw, dims[l]);
    for (int l = 0; l < n_layers; ++l) per_member += (int64_t)dims[l] * dims[l + 1] + dims[l + 1];
    // row batches sized so that the two activation buffers stay within 1 GiB
    const int64_t NB = std::max<int64_t>(1024, std::min<int64_t>(N, ((int64_t)1 << 30) / (16 * (int64_t)maxw)));
    PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)(2 * NB * maxw + NB * n_L + 64)));
    double* H[2] = {ctx->ws, ctx->ws + NB * maxw};
    double* var1 = ctx->ws + 2 * NB * maxw;
    auto blocks = [](int64_t n) { return (unsigned)((n + 255) / 256); };
    for (int64_t r0 = 0; r0 < N; r0 += NB) {
        const int64_t nr = std::min(NB, N - r0);
        double* s1 = v_mean + r0 * n_L; double* s2 = v_sig + r0 * n_L;
        for (int mem = 0; mem < M; ++mem) {
            const double* wm = w + (int64_t)mem * per_member;
            normalize_rows_kernel<<<blocks(nr * dims[0]), 256, 0, ctx->stream>>>(X + r0 * dims[0], nr, dims[0], norm_kind, na, nb, H[0]);
            PB_LAUNCH_CHECK(ctx);
            int cur = 0;
            for (int l = 0; l < n_layers; ++l) {
                const int in = dims[l], out = dims[l + 1];
                PB_TRY(pbi_gemm_nn(ctx, H[cur], in, nr, in, wm, out, out, H[cur ^ 1], out, nullptr, 0, nullptr));
                bias_act_kernel<<<blocks(nr * out), 256, 0, ctx->stream>>>(H[cur ^ 1], nr, out, wm + (int64_t)in * out, l + 1 < n_layers);
                PB_LAUNCH_CHECK(ctx);
                wm += (int64_t)in * out + out;
                cur ^= 1;
            }
            mixture_accum_kernel<<<blocks(nr * n_L), 256, 0, ctx->stream>>>(H[cur], nr, n_L, soft0, mem == 0, s1, s2, var1);
            PB_LAUNCH_CHECK(ctx);
        }
        mixture_final_kernel<<<blocks(nr * n_L), 256, 0, ctx->stream>>>(s1, s2, var1, nr * n_L, M);
        PB_LAUNCH_CHECK(ctx);
    }
    return PB_OK;
}

}  // namespace

extern "C" int pb_ensemble_predict(pb_ctx* ctx, int n_members, int n_layers, const int* dims,
                                   const double* weights_dev, int norm_kind, const double* norm_a_dev,
                                   const double* norm_b_dev, double soft_0, const double* X_dev, int64_t N,
                                   double* v_mean_dev, double* v_sig_dev) {
    PB_ENTER(ctx); if (!dims) return PB_ERR_ARG;
    if (n_members <= 0 || n_layers < 1 || n_layers > MAX_LAYERS || N <= 0)
        PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_predict: need 1..%d layers, >= 1 member, N > 0", MAX_LAYERS);
    const int out_last = dims[n_layers];
    if (out_last % 2) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_predict: the head must have 2*n_L outputs");
    bool wide = false;
    for (int l = 0; l < n_layers; ++l) {
        if (dims[l] <= 0 || dims[l] > 8192)
            PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_predict: layer input width %d outside (0, 8192]", dims[l]);
        wide = wide || dims[l] > MAX_WIDTH;
    }
    if (norm_kind != PB_NORM_NONE && (!norm_a_dev || !norm_b_dev)) PB_FAIL(ctx, PB_ERR_ARG, "pb_ensemble_predict: normalisation bounds missing");
    if (wide) {     // a layer wider than the fused kernel's activation tile: layer-by-layer path (any h_layers, varneuralnetwork.py:47-50)
        cudaEventRecord(ctx->ev0, ctx->stream);
        PB_TRY(ensemble_predict_layered(ctx, n_members, n_layers, dims, weights_dev, norm_kind, norm_a_dev, norm_b_dev, soft_0,
                                        X_dev, N, v_mean_dev, v_sig_dev));
        cudaEventRecord(ctx->ev1, ctx->stream);
        PB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->ms[PB_T_PREDICT] = ms;
        return PB_OK;
    }
    MlpArgs a{};
    a.n_members = n_members; a.n_layers = n_layers; a.n_L = out_last / 2; a.norm_kind = norm_kind;
    int64_t packed = 0, src = 0;
    for (int l = 0; l <= n_layers; ++l) a.dims[l] = dims[l];
    for (int l = 0; l < n_layers; ++l) {
        const int nk = (dims[l] + BK - 1) / BK, nc = (dims[l + 1] + 127) / 128;
        packed += (int64_t)nc * ((int64_t)nk * CHUNK_DOUBLES + BIAS_DOUBLES);
        src += (int64_t)dims[l] * dims[l + 1] + dims[l + 1];
    }
    a.member_stride = packed;
    a.norm_a = norm_a_dev; a.norm_b = norm_b_dev; a.soft0 = soft_0;
    a.X = X_dev; a.N = N; a.v_mean = v_mean_dev; a.v_sig = v_sig_dev;
    static const int groups = []{ const char* e = getenv("PB_MLP_GROUPS"); const int g = e ? atoi(e) : MLP_GROUPS; return (g >= 1 && g <= 3) ? g : MLP_GROUPS; }();
    const int MPT = 32 * groups;
    a.n_tiles = (N + MPT - 1) / MPT;
    { static const unsigned sk = []{ const char* e = getenv("PB_MLP_SKEW_NS"); return e ? (unsigned)atoi(e) : 3000u; }(); a.skew_ns = sk; }
    const int grid = (int)std::min<int64_t>(a.n_tiles, ctx->sm_count);
    cudaEventRecord(ctx->ev0, ctx->stream);
    // workspace: packed weights, then the per-CTA member scratch
    const size_t scr = (size_t)grid * n_members * MPT * 2 * a.n_L;
    PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)n_members * packed + scr + 16));
    double* pk = ctx->ws;
    a.packed = pk; a.scratch = pk + (((size_t)n_members * packed + 1) & ~(size_t)1);
    pack_weights_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(weights_dev, pk, n_members, n_layers, a, src);
    PB_LAUNCH_CHECK(ctx);
    const size_t smem = sizeof(double) * (size_t)(MPT * PACT + RING * CHUNK_DOUBLES + MPT * 4) + 2 * RING * sizeof(uint64_t);
    if (groups == 3) {
        PB_TRY(pbi_smem_optin(ctx, (const void*)ensemble_kernel<3>, (size_t)(smem)));
        ensemble_kernel<3><<<grid, (3 * 4 + 1) * 32, smem, ctx->stream>>>(a);
    } else if (groups == 2) {
        PB_TRY(pbi_smem_optin(ctx, (const void*)ensemble_kernel<2>, (size_t)(smem)));
        ensemble_kernel<2><<<grid, (2 * 4 + 1) * 32, smem, ctx->stream>>>(a);
    } else {
        PB_TRY(pbi_smem_optin(ctx, (const void*)ensemble_kernel<1>, (size_t)(smem)));
        ensemble_kernel<1><<<grid, (1 * 4 + 1) * 32, smem, ctx->stream>>>(a);
    }
    PB_LAUNCH_CHECK(ctx);
    cudaEventRecord(ctx->ev1, ctx->stream);
    PB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->ms[PB_T_PREDICT] = ms;
    return PB_OK;
}

extern "C" int pb_predict_U(pb_ctx* ctx, const double* V_dev, int64_t ldv, int64_t n_rows, int n_L,
                            const double* v_mean_dev, const double* v_sig_dev, int64_t N,
                            int samples, uint64_t seed, double* U_mean_dev, double* U_sig_dev, int64_t ldo) {
    PB_ENTER(ctx);
    if (n_L <= 0 || N <= 0 || n_rows <= 0 || ldo < N) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_predict_U: bad shape");
    if (samples < 0 || samples == 1) PB_FAIL(ctx, PB_ERR_ARG, "pb_predict_U: samples must be 0 (closed form) or >= 2");
    if (samples > 0) {
        // the reference's loop: S draws of v ~ N(v_mean, v_sig), U_i = V v_i^T, running sum and sum of squares
        if (!U_mean_dev || !U_sig_dev) PB_FAIL(ctx, PB_ERR_ARG, "pb_predict_U: the Monte-Carlo path needs both outputs");
        const int64_t ldb = (N + 1) / 2 * 2;
        PB_TRY(pb_reserve(ctx, &ctx->Bmat, &ctx->B_cap, (size_t)n_L * ldb));
        PB_CUDA(ctx, cudaMemset2DAsync(U_mean_dev, ldo * sizeof(double), 0, N * sizeof(double), n_rows, ctx->stream));
        PB_CUDA(ctx, cudaMemset2DAsync(U_sig_dev, ldo * sizeof(double), 0, N * sizeof(double), n_rows, ctx->stream));
        const int64_t pairs = (N * n_L + 1) / 2;
        for (int d = 0; d < samples; ++d) {
            sample_v_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, ctx->stream>>>(v_mean_dev, v_sig_dev, N, n_L, seed,
                                                                                  (uint32_t)d, ctx->Bmat, ldb);
            PB_LAUNCH_CHECK(ctx);
            // fused epilogue: U_i is never stored, the tile adds itself to the running moments
            PB_TRY(pbi_gemm_nn(ctx, V_dev, ldv, n_rows, n_L, ctx->Bmat, ldb, N, nullptr, 0, U_mean_dev, -ldo, U_sig_dev));
        }
        const int64_t n = n_rows * N;
        mc_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(U_mean_dev, U_sig_dev, n_rows, N, ldo, (double)samples);
        PB_LAUNCH_CHECK(ctx);
        return PB_OK;
    }
    // U_mean = V (n_rows x n_L) * v_mean^T (n_L x N)
    const int64_t ldb = (N + 1) / 2 * 2;
    PB_TRY(pb_reserve(ctx, &ctx->Bmat, &ctx->B_cap, (size_t)n_L * ldb));
    if (U_mean_dev) {
        PB_TRY(pbi_transpose(ctx, v_mean_dev, N, n_L, n_L, ctx->Bmat, ldb));
        PB_TRY(pbi_gemm_nn(ctx, V_dev, ldv, n_rows, n_L, ctx->Bmat, ldb, N, U_mean_dev, ldo, nullptr, 0, nullptr));
    }
    if (U_sig_dev) {
        // (V*V) (sig^2)^T, then sqrt
        // transient workspace of the context (grown on demand, kept): V*V (n_rows x n_L) then sig^2 (N x n_L)
        PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)n_rows * n_L + (size_t)N * n_L));
        double* V2 = ctx->ws; double* s2 = V2 + (size_t)n_rows * n_L;
        int64_t nv = n_rows * n_L, ns = N * n_L;
        square_strided_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, ctx->stream>>>(V_dev, ldv, V2, n_rows, n_L);
        PB_LAUNCH_CHECK(ctx);
        square_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, ctx->stream>>>(v_sig_dev, s2, ns);
        PB_LAUNCH_CHECK(ctx);
        PB_TRY(pbi_transpose(ctx, s2, N, n_L, n_L, ctx->Bmat, ldb));
        PB_TRY(pbi_gemm_nn(ctx, V2, n_L, n_rows, n_L, ctx->Bmat, ldb, N, U_sig_dev, ldo, nullptr, 0, nullptr));
        int64_t n = n_rows * N;
        sqrt_inplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(U_sig_dev, n_rows, N, ldo);
        PB_LAUNCH_CHECK(ctx);
    }
    return PB_OK;
}

extern "C" int pb_mixture_accumulate(pb_ctx* ctx, const double* U_mean_dev, const double* U_sig_dev, int64_t ld_in,
                                     int64_t n_rows, int64_t N, double* sum_dev, double* m2_dev, int64_t ld_acc, int first) {
    PB_ENTER(ctx);
    if (n_rows <= 0 || N <= 0 || ld_in < N || ld_acc < N || !U_mean_dev || !U_sig_dev || !sum_dev || !m2_dev)
        PB_FAIL(ctx, PB_ERR_SHAPE, "pb_mixture_accumulate: bad shape");
    const int64_t n = n_rows * N;
    u_mixture_accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(U_mean_dev, U_sig_dev, ld_in, sum_dev, m2_dev,
                                                                                   ld_acc, n_rows, N, first);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

extern "C" int pb_mixture_finalize(pb_ctx* ctx, double* sum_dev, double* m2_dev, int64_t ld, int64_t n_rows, int64_t N,
                                   int n_members, const double* pod_sig_dev) {
    PB_ENTER(ctx);
    if (n_rows <= 0 || N <= 0 || ld < N || n_members <= 0 || !sum_dev || !m2_dev)
        PB_FAIL(ctx, PB_ERR_SHAPE, "pb_mixture_finalize: bad shape");
    const int64_t n = n_rows * N;
    u_mixture_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(sum_dev, m2_dev, ld, n_rows, N, (double)n_members,
                                                                                 pod_sig_dev);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}
