// Deep-ensemble training step on the fp64 tensor pipe.
//
// Replaces VarNeuralNetwork.grad / tf_optimization_step / tf_optimization
// (varneuralnetwork.py:94-128): per epoch and member
//   loss = NLL(X) + reg                       NLL = sum -Normal(loc, softplus(soft_0 t_s) + 1e-6).log_prob(v)
//   X_adv = X + adv_eps sign(d loss / d X)    (when adv_eps is not None; no gradient through sign)
//   loss += NLL(X_adv) + reg                  reg = lam * sum_k sum(var_k^2)/2, kernels and biases
//   Keras Adam (beta 0.9 / 0.999, eps 1e-7) on d loss / d vars, full batch.
//
// All M members train in the same launches (batch index = blockIdx.z).  One epoch is a fixed
// sequence of ~26 small kernels: forward Dense layers (bias/ReLU epilogues), the NLL + head
// gradient, the delta chain (ReLU-mask epilogue, FGSM sign epilogue on the input gradient), the
// weight gradients H^T delta as split-K partials (bias gradient = column sums from the same
// tiles), and one Adam kernel that sums the partials in fixed order, adds the L2 term, updates
// the parameters and their transposes and assembles the loss.  The sequence is captured once in
// a CUDA graph (weight-gradient products on a forked branch, off the critical path) and replayed
// per epoch: at the reference's sizes (N = 400..4000 rows, 128-wide layers) the step is bound by
// kernel latency, not by flops.
#include "common.cuh"
#include <mutex>
#include "tile_ops.cuh"

namespace {
using namespace pbt;

constexpr int MAXL = 8;
constexpr int STAGES = 4;
constexpr int NLL_BLOCKS_MAX = 256;
constexpr int MAX_SPLITS = 24;

struct LayerTab {
    int n_layers;
    int in[MAXL], out[MAXL];
    long long off_w[MAXL], off_b[MAXL], off_wt[MAXL];
    long long P, PT;
    // fused path: packed weight streams (see train_fused_kernel)
    int fused;
    long long pk_f[MAXL], pk_b[MAXL], SP;
};

constexpr int FPITCH = 132;                 // pitch of the staged activations and of the packed weight rows
constexpr int FCH = BK * FPITCH;            // one packed K chunk: 16 rows x 132 doubles = 16 896 B
constexpr int FRING = 4;

// where parameter p of a member lives in the derived layouts: the transposed kernels (layer-by-layer path) or the
// packed forward / backward streams (fused path)
__device__ __forceinline__ void scatter_param(const LayerTab& tab, double* __restrict__ WT, double* __restrict__ packed,
                                              int mem, long long p, double val) {
    for (int l = 0; l < tab.n_layers; ++l) {
        const long long q = p - tab.off_w[l];
        if (q >= 0 && q < (long long)tab.in[l] * tab.out[l]) {
            const int i = (int)(q / tab.out[l]), j = (int)(q % tab.out[l]);
            if (tab.fused) {
                double* pk = packed + (long long)mem * tab.SP;
                pk[tab.pk_f[l] + (long long)(i / BK) * FCH + (i % BK) * FPITCH + j] = val;
                pk[tab.pk_b[l] + (long long)(j / BK) * FCH + (j % BK) * FPITCH + i] = val;
            } else {
                WT[(long long)mem * tab.PT + tab.off_wt[l] + (long long)j * tab.in[l] + i] = val;
            }
            return;
        }
        const long long qb = p - tab.off_b[l];
        if (tab.fused && qb >= 0 && qb < tab.out[l]) {
            const int nk = (tab.in[l] + BK - 1) / BK;
            packed[(long long)mem * tab.SP + tab.pk_f[l] + (long long)nk * FCH + qb] = val;
            return;
        }
    }
}

enum { EPI_BIAS_RELU = 0, EPI_BIAS = 1, EPI_MASK = 2, EPI_SIGN = 3 };

template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
// (programmatic dependent launch between the per-layer kernels was measured and rejected: profiles/r01_notes.md)

// ---------------------------------------------------------------------------
// Y (rows x k) = A (rows x m) B (m x k) with a fused epilogue; blockIdx.z = member
// ---------------------------------------------------------------------------
template <int BM, int BN, int WARPS_M, int WARPS_N, bool ALIGN_A, bool ALIGN_B>
__global__ void __launch_bounds__(NTHREADS)
train_nn_kernel(const double* __restrict__ A, int64_t lda, int64_t sA, int64_t rows, int m,
                const double* __restrict__ B, int64_t ldb, int64_t sB, int k,
                double* __restrict__ Y, int64_t ldy, int64_t sY,
                int epi, const double* __restrict__ aux, int64_t ldaux, int64_t sAux, double eps) {
    static_assert(WARPS_M * WARPS_N == NTHREADS / 32, "8 warps");
    constexpr int PA = BK + 4, PB = BN + 4;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int TM = WM / 8, TNN = WN / 8;
    A += (int64_t)blockIdx.z * sA; B += (int64_t)blockIdx.z * sB; Y += (int64_t)blockIdx.z * sY;
    aux += (int64_t)blockIdx.z * sAux;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                              // [STAGES][BM][PA]
    double* Bs = smem + STAGES * BM * PA;           // [STAGES][BK][PB]

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

#pragma unroll
    for (int t = 0; t < TM; ++t) {
        const int64_t h = h0 + wm * WM + t * 8 + lg;
        if (h >= rows) continue;
#pragma unroll
        for (int u = 0; u < TNN; ++u) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t j = j0 + wn * WN + u * 8 + 2 * lk + e;
                if (j >= k) continue;
                double x = acc[t][u][e];
                if (epi == EPI_BIAS_RELU)      x = fmax(x + aux[j], 0.);
                else if (epi == EPI_BIAS)      x = x + aux[j];
                else if (epi == EPI_MASK)      x = (aux[h * ldaux + j] > 0.) ? x : 0.;
                else /* EPI_SIGN */            x = aux[h * ldaux + j] + eps * ((x > 0.) ? 1. : ((x < 0.) ? -1. : 0.));
                Y[h * ldy + j] = x;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Weight-gradient partials: out (ka x kb) = A^T B over the row range of split blockIdx.y, and (tile row 0 only)
// the column sums of B over the same range = bias-gradient partial.  blockIdx.z = member.
// ---------------------------------------------------------------------------
template <int BM, int BN, int WARPS_M, int WARPS_N, bool ALIGN_A, bool ALIGN_B>
__global__ void __launch_bounds__(NTHREADS)
train_dw_kernel(const double* __restrict__ A, int64_t lda, int64_t sA, int ka,
                const double* __restrict__ B, int64_t ldb, int64_t sB, int kb,
                int64_t rows, int64_t rows_per_split,
                double* __restrict__ out, int64_t ldo, int64_t split_stride, int64_t sOut,
                double* __restrict__ bias_out, int tiles_n) {
    static_assert(WARPS_M * WARPS_N == NTHREADS / 32, "8 warps");
    constexpr int PA = BM + 4, PB = BN + 4;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int TM = WM / 8, TNN = WN / 8;
    A += (int64_t)blockIdx.z * sA; B += (int64_t)blockIdx.z * sB;
    out += (int64_t)blockIdx.z * sOut + (int64_t)blockIdx.y * split_stride;
    bias_out += (int64_t)blockIdx.z * sOut + (int64_t)blockIdx.y * split_stride;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                              // [STAGES][BK][PA]
    double* Bs = smem + STAGES * BK * PA;           // [STAGES][BK][PB]

    const int ti = blockIdx.x / tiles_n, tj = blockIdx.x % tiles_n;
    const int64_t i0 = (int64_t)ti * BM, j0 = (int64_t)tj * BN;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_end = min(rows, r_begin + rows_per_split);
    const int nk = (r_end > r_begin) ? (int)((r_end - r_begin + BK - 1) / BK) : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int lk = lane & 3, lg = lane >> 2;

    double acc[TM][TNN][2];
#pragma unroll
    for (int t = 0; t < TM; ++t)
#pragma unroll
        for (int u = 0; u < TNN; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }
    double csum = 0.;
    const bool do_bias = (ti == 0) && tid < BN;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            int64_t r0 = r_begin + (int64_t)s * BK;
            load_tile<BK, BM, PA, ALIGN_A>(As + s * BK * PA, A, lda, r0, r_end, i0, ka, tid);
            load_tile<BK, BN, PB, ALIGN_B>(Bs + s * BK * PB, B, ldb, r0, r_end, j0, kb, tid);
        }
        cp_async_commit();
    }
    for (int k = 0; k < nk; ++k) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int kn = k + STAGES - 1;
            if (kn < nk) {
                int s = kn % STAGES;
                int64_t r0 = r_begin + (int64_t)kn * BK;
                load_tile<BK, BM, PA, ALIGN_A>(As + s * BK * PA, A, lda, r0, r_end, i0, ka, tid);
                load_tile<BK, BN, PB, ALIGN_B>(Bs + s * BK * PB, B, ldb, r0, r_end, j0, kb, tid);
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
        if (do_bias) {
            const double* bc = Bs + (k % STAGES) * BK * PB + tid;
#pragma unroll
            for (int r = 0; r < BK; ++r) csum += bc[r * PB];      // rows beyond r_end are zero-filled
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int t = 0; t < TM; ++t) {
        int64_t i = i0 + wm * WM + t * 8 + lg;
        if (i >= ka) continue;
#pragma unroll
        for (int u = 0; u < TNN; ++u) {
            int64_t j = j0 + wn * WN + u * 8 + 2 * lk;
            if (j < kb)     out[i * ldo + j]     = acc[t][u][0];
            if (j + 1 < kb) out[i * ldo + j + 1] = acc[t][u][1];
        }
    }
    if (do_bias && j0 + tid < kb) bias_out[j0 + tid] = csum;
}

// ---------------------------------------------------------------------------
// NLL and head gradient, in place: T (N x 2 n_L) -> d loss / d T.  grid (NLL_BLOCKS, M)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
train_nll_kernel(double* __restrict__ T, int64_t sT, const double* __restrict__ v, int64_t N, int n_L,
                 double soft_0, double* __restrict__ nllpart) {
    T += (int64_t)blockIdx.y * sT;
    const int64_t total = N * n_L;
    double s = 0.;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = idx / n_L; const int j = (int)(idx % n_L);
        double* row = T + n * 2 * n_L;
        const double loc = row[j], z = soft_0 * row[n_L + j];
        const double e = exp(-fabs(z));
        const double scale = fmax(z, 0.) + log1p(e) + 1e-6;
        const double inv = 1. / scale;
        const double r = (v[idx] - loc) * inv;
        s += 0.5 * r * r + log(scale) + 0.91893853320467274178;
        const double sg = (z >= 0.) ? 1. / (1. + e) : e / (1. + e);
        row[j] = -r * inv;
        row[n_L + j] = (1. - r * r) * inv * soft_0 * sg;
    }
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int w = 0; w < 8; ++w) t += red[w];
        nllpart[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------
// Adam: g = sum of the split-K partials of both passes (fixed order) + n_reg lam w; Keras update; transposed copy of
// the kernels for the next delta chain; loss assembly by the last block of each member.  grid (ceil(P/256), M)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
train_adam_kernel(LayerTab tab, double* __restrict__ params, double* __restrict__ am, double* __restrict__ av,
                  double* __restrict__ WT, double* __restrict__ packed, const double* __restrict__ gpart, int n_pass,
                  int n_term, int splits, int M,
                  double lr, double lam, long long* __restrict__ tstep, int* __restrict__ eidx,
                  const double* __restrict__ nllpart, int nll_blocks, double* __restrict__ regpart,
                  unsigned* __restrict__ ticket, double* __restrict__ loss, int update) {
    const int mem = blockIdx.y;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ double s_lrt;
    __shared__ double red[8];
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        const double t = (double)(tstep[mem] + 1);
        s_lrt = lr * sqrt(1. - pow(0.999, t)) / (1. - pow(0.9, t));
    }
    __syncthreads();
    double w2 = 0.;
    if (p < tab.P) {
        double* w = params + (long long)mem * tab.P + p;
        const double w0 = *w;
        w2 = w0 * w0;
        if (update) {
            double g = 0.;
            for (int q = 0; q < n_pass * splits; ++q) g += gpart[((long long)q * M + mem) * tab.P + p];
            // n_term = NLL + L2 terms in the loss (2 with the adversarial branch); when adv_eps == 0 the second
            // term equals the first exactly (X_adv = X + 0 sign(.) = X), so one pass is run and counted twice
            g = g * (double)(n_term / n_pass) + (double)n_term * lam * w0;
            double* mp = am + (long long)mem * tab.P + p; double* vp = av + (long long)mem * tab.P + p;
            const double mn = 0.9 * (*mp) + (1. - 0.9) * g;
            const double vn = 0.999 * (*vp) + (1. - 0.999) * g * g;
            *mp = mn; *vp = vn;
            const double wn = w0 - s_lrt * mn / (sqrt(vn) + 1e-7);
            *w = wn;
            scatter_param(tab, WT, packed, mem, p, wn);
        } else {
            scatter_param(tab, WT, packed, mem, p, w0);
        }
    }
    if (!update) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w2 += __shfl_xor_sync(0xffffffffu, w2, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int w = 0; w < 8; ++w) t += red[w];
        regpart[(long long)mem * gridDim.x + blockIdx.x] = t;
        __threadfence();
        const unsigned done = atomicAdd(&ticket[mem], 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {      // fixed-order block reduction of the partial sums (deterministic for a given launch shape)
        __threadfence();
        double reg = 0., nll = 0.;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
            reg += ((volatile double*)regpart)[(long long)mem * gridDim.x + b];
        for (int q = 0; q < n_pass; ++q)
            for (int b = threadIdx.x; b < nll_blocks; b += blockDim.x)
                nll += nllpart[((long long)q * M + mem) * nll_blocks + b];
        double tot = nll * (double)(n_term / n_pass) + (double)n_term * lam * 0.5 * reg;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.;
            for (int w = 0; w < 8; ++w) t += red[w];
            loss[(long long)eidx[mem] * M + mem] = t;
            eidx[mem] += 1;
            tstep[mem] += 1;
            ticket[mem] = 0;
        }
    }
}


// ---------------------------------------------------------------------------
// Fused forward + NLL + delta chain (+ FGSM input and the second pass) for one (member, row tile): the activations
// stay in shared memory from the input to the input gradient, the ReLU masks stay in registers, and the weights
// arrive as a continuous packed stream (forward kernels + bias rows, then the transposed kernels in reverse layer
// order) through a 4-slot cp.async.bulk / mbarrier ring filled by a producer warp that runs ahead across layer,
// pass and tile boundaries.  Only what the weight-gradient products need goes to global memory: H_l, delta_l and
// X_adv.  Requires every width <= 128 (one 128-column chunk per layer); wider nets take the layer-by-layer path.
// ---------------------------------------------------------------------------
struct FusedArgs {
    int M, L, n_L, n_pass, n_tiles;
    long long N;
    int dims[MAXL + 1];
    const double* packed; long long SP; long long pk_f[MAXL], pk_b[MAXL];
    const double* X; const double* v;
    double* xadv;
    double* act[2]; double* del[2];
    long long act_off[MAXL + 1], del_off[MAXL + 1];
    double* nllpart;
    double adv_eps, soft_0;
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
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

constexpr int FUSED_THREADS = 288;          // 8 consumer warps + 1 producer warp

template <int MT>
__global__ void __launch_bounds__(FUSED_THREADS, 2) train_fused_kernel(FusedArgs a) {
    constexpr int TM = MT / 8;              // m8 row blocks per warp (every warp covers all MT rows, 16 columns)
    extern __shared__ __align__(16) double smem[];
    double* act = smem;                                   // [MT][FPITCH]
    double* ring = smem + MT * FPITCH;                    // [FRING][FCH]
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + FRING * FCH);
    uint64_t* empty = full + FRING;
    double* red = reinterpret_cast<double*>(empty + FRING);   // [8]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < FRING; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    const int L = a.L, n_L = a.n_L;
    const long long n_items = (long long)a.M * a.n_tiles;

    if (warp == 8) {
        // ------------------------------------------------ producer
        if (lane == 0) {
            uint32_t it = 0;
            for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int mem = (int)(item / a.n_tiles);
                const double* base = a.packed + (long long)mem * a.SP;
                for (int pass = 0; pass < a.n_pass; ++pass) {
                    for (int l = 0; l < L; ++l) {                       // forward kernels, each followed by its bias row
                        const int nk = (a.dims[l] + BK - 1) / BK;
                        const double* src = base + a.pk_f[l];
                        for (int kc = 0; kc <= nk; ++kc, ++it) {
                            const int s = it % FRING; const uint32_t ph = (it / FRING) & 1;
                            const uint32_t bytes = (kc < nk ? FCH : FPITCH) * 8;
                            mbar_wait(empty + s, ph ^ 1);
                            mbar_expect_tx(full + s, bytes);
                            bulk_load(ring + (size_t)s * FCH, src + (size_t)kc * FCH, bytes, full + s);
                        }
                    }
                    const int l_last = (pass == 0 && a.n_pass == 2) ? 0 : 1;
                    for (int l = L - 1; l >= l_last; --l) {             // transposed kernels for the delta chain
                        const int nk = (a.dims[l + 1] + BK - 1) / BK;
                        const double* src = base + a.pk_b[l];
                        for (int kc = 0; kc < nk; ++kc, ++it) {
                            const int s = it % FRING; const uint32_t ph = (it / FRING) & 1;
                            mbar_wait(empty + s, ph ^ 1);
                            mbar_expect_tx(full + s, FCH * 8);
                            bulk_load(ring + (size_t)s * FCH, src + (size_t)kc * FCH, FCH * 8, full + s);
                        }
                    }
                }
            }
        }
        return;
    }

    // ---------------------------------------------------- consumers: warp w owns columns [16 w, 16 w + 16)
    const int lk = lane & 3, lg = lane >> 2;
    const int n_d = a.dims[0];
    uint32_t it = 0;
    double acc[TM][2][2];

    // acc = act(:, 0:16 nk) * (nk packed chunks from the ring); warps whose 16 columns lie beyond `ncols` only keep the
    // ring moving (narrow outputs: the head, the input gradient)
    auto gemm = [&](int nk, int ncols) {
        const bool active = warp * 16 < ncols;
#pragma unroll
        for (int t = 0; t < TM; ++t)
#pragma unroll
            for (int u = 0; u < 2; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }
#pragma unroll 1
        for (int kt = 0; kt < nk; ++kt, ++it) {
            const int s = it % FRING; const uint32_t ph = (it / FRING) & 1;
            mbar_wait(full + s, ph);
            const double* as = act + lg * FPITCH + kt * BK + lk;
            const double* bs = ring + (size_t)s * FCH + warp * 16 + lg;
            if (active)
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                double af[TM], bf[2];
#pragma unroll
                for (int t = 0; t < TM; ++t) af[t] = as[t * 8 * FPITCH + kk * 4];
#pragma unroll
                for (int u = 0; u < 2; ++u) bf[u] = bs[(kk * 4 + lk) * FPITCH + u * 8];
#pragma unroll
                for (int t = 0; t < TM; ++t)
#pragma unroll
                    for (int u = 0; u < 2; ++u) dmma884(acc[t][u][0], acc[t][u][1], af[t], bf[u]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
    };

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mem = (int)(item / a.n_tiles);
        const int tile = (int)(item % a.n_tiles);
        const long long r0 = (long long)tile * MT;
        const int kin0 = (n_d + BK - 1) / BK * BK;
        // input tile, zero padded to 16 columns and to MT rows
        for (int e = tid; e < MT * kin0; e += 256) {
            const int r = e / kin0, c = e % kin0;
            act[r * FPITCH + c] = (c < n_d && r0 + r < a.N) ? a.X[(r0 + r) * n_d + c] : 0.;
        }
        consumer_sync();
        for (int pass = 0; pass < a.n_pass; ++pass) {
            uint32_t mask[MAXL];               // ReLU masks of the hidden layers, one bit per accumulator element
            // ---------------- forward
            for (int l = 0; l < L; ++l) {
                const int in = a.dims[l], out = a.dims[l + 1];
                const bool head = (l == L - 1);
                gemm((in + BK - 1) / BK, out);
                const int sb = it % FRING; const uint32_t phb = (it / FRING) & 1;
                ++it;
                mbar_wait(full + sb, phb);
                const double* bias = ring + (size_t)sb * FCH;
                consumer_sync();               // every warp has finished reading act for this layer
                const int outp = (out + BK - 1) / BK * BK;
                double* gout = head ? nullptr : a.act[pass] + a.act_off[l + 1] + ((long long)mem * a.N + r0) * out;
                const bool vec_ok = !(out & 1) && ((reinterpret_cast<uintptr_t>(gout) & 15) == 0);
                uint32_t mk = 0;
#pragma unroll
                for (int t = 0; t < TM; ++t) {
                    const int r = t * 8 + lg;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int n = warp * 16 + u * 8 + 2 * lk;
                        double x0 = acc[t][u][0] + bias[n], x1 = acc[t][u][1] + bias[n + 1];
                        if (!head) {
                            x0 = fmax(x0, 0.); x1 = fmax(x1, 0.);
                            mk |= (x0 > 0. ? 1u : 0u) << (t * 4 + u * 2);
                            mk |= (x1 > 0. ? 1u : 0u) << (t * 4 + u * 2 + 1);
                        }
                        if (n >= out) x0 = 0.;
                        if (n + 1 >= out) x1 = 0.;
                        if (n < outp) *reinterpret_cast<double2*>(act + r * FPITCH + n) = make_double2(x0, x1);
                        if (!head && r0 + r < a.N) {
                            if (n + 1 < out && vec_ok) *reinterpret_cast<double2*>(gout + (long long)r * out + n) = make_double2(x0, x1);
                            else { if (n < out) gout[(long long)r * out + n] = x0; if (n + 1 < out) gout[(long long)r * out + n + 1] = x1; }
                        }
                    }
                }
                if (!head) mask[l] = mk;
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + sb);
                consumer_sync();               // the next layer's input is complete
            }
            // ---------------- NLL and head gradient, in place on the staged head outputs
            {
                const int w2 = 2 * n_L;
                double* gd = a.del[pass] + a.del_off[L - 1] + ((long long)mem * a.N + r0) * w2;
                double sacc = 0.;
                for (int e = tid; e < MT * n_L; e += 256) {
                    const int r = e / n_L, j = e % n_L;
                    double* row = act + r * FPITCH;
                    double d0 = 0., d1 = 0.;
                    if (r0 + r < a.N) {
                        const double loc = row[j], z = a.soft_0 * row[n_L + j];
                        const double ex = exp(-fabs(z));
                        const double scale = fmax(z, 0.) + log1p(ex) + 1e-6;
                        const double inv = 1. / scale;
                        const double rr = (a.v[(r0 + r) * n_L + j] - loc) * inv;
                        sacc += 0.5 * rr * rr + log(scale) + 0.91893853320467274178;
                        const double sg = (z >= 0.) ? 1. / (1. + ex) : ex / (1. + ex);
                        d0 = -rr * inv;
                        d1 = (1. - rr * rr) * inv * a.soft_0 * sg;
                        gd[(long long)r * w2 + j] = d0;
                        gd[(long long)r * w2 + n_L + j] = d1;
                    }
                    row[j] = d0; row[n_L + j] = d1;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
                if (lane == 0) red[warp] = sacc;
                consumer_sync();
                if (tid == 0) {
                    double tsum = 0.;
                    for (int w = 0; w < 8; ++w) tsum += red[w];
                    a.nllpart[((long long)pass * a.M + mem) * a.n_tiles + tile] = tsum;
                }
            }
            // ---------------- delta chain (and the FGSM input after the first pass)
            const int l_last = (pass == 0 && a.n_pass == 2) ? 0 : 1;
            for (int l = L - 1; l >= l_last; --l) {
                const int in = a.dims[l], out = a.dims[l + 1];     // delta_l (MT x out) -> delta_{l-1} (MT x in)
                gemm((out + BK - 1) / BK, in);
                consumer_sync();
                const int inp = (in + BK - 1) / BK * BK;
                if (l > 0) {
                    double* gout = a.del[pass] + a.del_off[l - 1] + ((long long)mem * a.N + r0) * in;
                    const uint32_t mk = mask[l - 1];
                    const bool vec_ok = !(in & 1) && ((reinterpret_cast<uintptr_t>(gout) & 15) == 0);
#pragma unroll
                    for (int t = 0; t < TM; ++t) {
                        const int r = t * 8 + lg;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int n = warp * 16 + u * 8 + 2 * lk;
                            double x0 = ((mk >> (t * 4 + u * 2)) & 1u) ? acc[t][u][0] : 0.;
                            double x1 = ((mk >> (t * 4 + u * 2 + 1)) & 1u) ? acc[t][u][1] : 0.;
                            if (n >= in) x0 = 0.;
                            if (n + 1 >= in) x1 = 0.;
                            if (n < inp) *reinterpret_cast<double2*>(act + r * FPITCH + n) = make_double2(x0, x1);
                            if (r0 + r < a.N) {
                                if (n + 1 < in && vec_ok) *reinterpret_cast<double2*>(gout + (long long)r * in + n) = make_double2(x0, x1);
                                else { if (n < in) gout[(long long)r * in + n] = x0; if (n + 1 < in) gout[(long long)r * in + n + 1] = x1; }
                            }
                        }
                    }
                } else {
                    // X_adv = X + adv_eps sign(d loss / d X): the next pass's input, and the A operand of its dW_0
                    double* gx = a.xadv + ((long long)mem * a.N + r0) * n_d;
#pragma unroll
                    for (int t = 0; t < TM; ++t) {
                        const int r = t * 8 + lg;
#pragma unroll
                        for (int u = 0; u < 2; ++u)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int n = warp * 16 + u * 8 + 2 * lk + e;
                                if (n >= inp) continue;
                                double x = 0.;
                                if (n < n_d && r0 + r < a.N) {
                                    const double g = acc[t][u][e];
                                    x = a.X[(r0 + r) * n_d + n] + a.adv_eps * ((g > 0.) ? 1. : ((g < 0.) ? -1. : 0.));
                                    gx[(long long)r * n_d + n] = x;
                                }
                                act[r * FPITCH + n] = x;
                            }
                    }
                }
                consumer_sync();
            }
        }
        consumer_sync();       // act is overwritten by the next item's input tile
    }
}

// ---------------------------------------------------------------------------
// All weight-gradient products of an epoch in one launch: jobs = (pass, layer), each cut into tiles and split-K
// row ranges like train_dw_kernel.  blockIdx.x = (job, tile), blockIdx.y = split, blockIdx.z = member.
// ---------------------------------------------------------------------------
struct DwJob {
    const double* A; const double* B;
    long long lda, sA, ldb, sB, out_off, bias_off;
    int ka, kb, tiles_n, tile0, alignA, alignB;
};
struct DwJobs { int n_jobs; DwJob job[2 * MAXL]; };

template <int BM, int BN, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(NTHREADS)
train_dw_multi_kernel(DwJobs jobs, int64_t rows, int64_t rows_per_split, double* __restrict__ gpart,
                      int64_t split_stride, int64_t sOut) {
    constexpr int PA = BM + 4, PB = BN + 4;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int TM = WM / 8, TNN = WN / 8;
    int jb = 0;
    while (jb + 1 < jobs.n_jobs && (int)blockIdx.x >= jobs.job[jb + 1].tile0) ++jb;
    const DwJob& J = jobs.job[jb];
    const double* A = J.A + (int64_t)blockIdx.z * J.sA;
    const double* B = J.B + (int64_t)blockIdx.z * J.sB;
    double* base = gpart + (int64_t)blockIdx.z * sOut + (int64_t)blockIdx.y * split_stride;
    double* out = base + J.out_off;
    double* bias_out = base + J.bias_off;
    const int ka = J.ka, kb = J.kb;
    const int64_t lda = J.lda, ldb = J.ldb, ldo = kb;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * BK * PA;
    const int tl = blockIdx.x - J.tile0;
    const int ti = tl / J.tiles_n, tj = tl % J.tiles_n;
    const int64_t i0 = (int64_t)ti * BM, j0 = (int64_t)tj * BN;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_end = min(rows, r_begin + rows_per_split);
    const int nk = (r_end > r_begin) ? (int)((r_end - r_begin + BK - 1) / BK) : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int lk = lane & 3, lg = lane >> 2;
    const bool alA = J.alignA != 0, alB = J.alignB != 0;
    // m8 / n8 blocks of this warp that hold valid rows / columns (narrow layers: the input layer has ka = n_d rows)
    const int t_cnt = max(0, min(TM, (int)((ka - i0 - wm * WM + 7) / 8)));
    const int u_cnt = max(0, min(TNN, (int)((kb - j0 - wn * WN + 7) / 8)));
    const bool full_tile = (t_cnt == TM) && (u_cnt == TNN), any_tile = (t_cnt > 0) && (u_cnt > 0);

    double acc[TM][TNN][2];
#pragma unroll
    for (int t = 0; t < TM; ++t)
#pragma unroll
        for (int u = 0; u < TNN; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }
    double csum = 0.;
    const bool do_bias = (ti == 0) && tid < BN;

    auto stage = [&](int s, int64_t r0) {
        if (alA) load_tile<BK, BM, PA, true>(As + s * BK * PA, A, lda, r0, r_end, i0, ka, tid);
        else     load_tile<BK, BM, PA, false>(As + s * BK * PA, A, lda, r0, r_end, i0, ka, tid);
        if (alB) load_tile<BK, BN, PB, true>(Bs + s * BK * PB, B, ldb, r0, r_end, j0, kb, tid);
        else     load_tile<BK, BN, PB, false>(Bs + s * BK * PB, B, ldb, r0, r_end, j0, kb, tid);
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) stage(s, r_begin + (int64_t)s * BK);
        cp_async_commit();
    }
    for (int k = 0; k < nk; ++k) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int kn = k + STAGES - 1;
            if (kn < nk) stage(kn % STAGES, r_begin + (int64_t)kn * BK);
            cp_async_commit();
        }
        const double* as = As + (k % STAGES) * BK * PA + wm * WM + lg;
        const double* bs = Bs + (k % STAGES) * BK * PB + wn * WN + lg;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double av[TM], bv[TNN];
#pragma unroll
            for (int t = 0; t < TM; ++t) av[t] = as[(kk * 4 + lk) * PA + t * 8];
#pragma unroll
            for (int u = 0; u < TNN; ++u) bv[u] = bs[(kk * 4 + lk) * PB + u * 8];
            if (full_tile) {
#pragma unroll
                for (int t = 0; t < TM; ++t)
#pragma unroll
                    for (int u = 0; u < TNN; ++u) dmma884(acc[t][u][0], acc[t][u][1], av[t], bv[u]);
            } else if (any_tile) {
#pragma unroll
                for (int t = 0; t < TM; ++t)
                    if (t < t_cnt)
#pragma unroll
                        for (int u = 0; u < TNN; ++u)
                            if (u < u_cnt) dmma884(acc[t][u][0], acc[t][u][1], av[t], bv[u]);
            }
        }
        if (do_bias) {
            const double* bc = Bs + (k % STAGES) * BK * PB + tid;
#pragma unroll
            for (int r = 0; r < BK; ++r) csum += bc[r * PB];
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int t = 0; t < TM; ++t) {
        int64_t i = i0 + wm * WM + t * 8 + lg;
        if (i >= ka) continue;
#pragma unroll
        for (int u = 0; u < TNN; ++u) {
            int64_t j = j0 + wn * WN + u * 8 + 2 * lk;
            if (j < kb)     out[i * ldo + j]     = acc[t][u][0];
            if (j + 1 < kb) out[i * ldo + j + 1] = acc[t][u][1];
        }
    }
    if (do_bias && j0 + tid < kb) bias_out[j0 + tid] = csum;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct TrainKey {
    int M, n_layers, dims[MAXL + 1], use_adv;
    int n_pass;          // forward/backward passes actually run: 2 only when adv_eps != 0
    int64_t N;
    const double *X, *v;
    double *params, *am, *av;
    double lr, lam, adv_eps, soft_0;
    bool operator==(const TrainKey& o) const { return memcmp(this, &o, sizeof(TrainKey)) == 0; }
};

struct TrainState {
    TrainKey key;
    bool valid = false;
    LayerTab tab;
    int splits = 1; int64_t rps = 0; int nll_blocks = 1; bool dw_big = false;
    bool fused = false; int mt = 64, n_tiles = 0, fused_grid = 0;       // fused forward/backward path
    double* packed = nullptr; size_t cap_packed = 0;
    DwJobs jobs; int dw_tiles = 0;
    int64_t act_off[MAXL + 1], del_off[MAXL + 1], act_total = 0, del_total = 0;
    double* act[2] = {nullptr, nullptr}; double* del[2] = {nullptr, nullptr};
    double *xadv = nullptr, *WT = nullptr, *gpart = nullptr, *nllpart = nullptr, *regpart = nullptr, *loss = nullptr;
    size_t cap_act[2] = {0, 0}, cap_del[2] = {0, 0}, cap_xadv = 0, cap_WT = 0, cap_gpart = 0, cap_nll = 0,
           cap_reg = 0, cap_loss = 0;
    unsigned* ticket = nullptr; long long* tstep = nullptr; int* eidx = nullptr;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaGraphExec_t exec = nullptr;
    int launches_per_epoch = 0;
};

inline bool al16(const void* p, int64_t ld, int64_t stride) {
    return ((uintptr_t)p % 16 == 0) && (ld % 2 == 0) && (stride % 2 == 0);
}

template <typename K>
int set_smem(pb_ctx* ctx, K kernel, size_t bytes) {
    PB_TRY(pbi_smem_optin(ctx, (const void*)kernel, (size_t)(bytes)));
    return PB_OK;
}

template <int BM, int BN, int WARPS_M, int WARPS_N, bool AA, bool AB>
int launch_nn_t(pb_ctx* ctx, cudaStream_t st, int M, const double* A, int64_t lda, int64_t sA, int64_t rows, int m,
                const double* B, int64_t ldb, int64_t sB, int k, double* Y, int64_t ldy, int64_t sY, int epi,
                const double* aux, int64_t ldaux, int64_t sAux, double eps) {
    auto kern = train_nn_kernel<BM, BN, WARPS_M, WARPS_N, AA, AB>;
    constexpr size_t smem = (size_t)STAGES * (BM * (BK + 4) + BK * (BN + 4)) * sizeof(double);
    if (!ctx) return set_smem(ctx, kern, smem);      // configuration pass (before any capture)
    dim3 grid((unsigned)((rows + BM - 1) / BM), (unsigned)((k + BN - 1) / BN), (unsigned)M);
    PB_CUDA(ctx, launch_k(kern, grid, dim3(NTHREADS), smem, st, A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps));
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

int launch_nn(pb_ctx* ctx, cudaStream_t st, int M, const double* A, int64_t lda, int64_t sA, int64_t rows, int m,
              const double* B, int64_t ldb, int64_t sB, int k, double* Y, int64_t ldy, int64_t sY, int epi,
              const double* aux, int64_t ldaux, int64_t sAux, double eps) {
    const bool aa = al16(A, lda, sA), ab = al16(B, ldb, sB);
    // 32-row tiles, 3 CTAs per SM: measured faster than 128-row tiles at every batch size (profiles/r01_notes.md)
#define PB_NN(BM_, BN_, WM_, WN_) \
    (aa ? (ab ? launch_nn_t<BM_, BN_, WM_, WN_, true, true>(ctx, st, M, A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps) \
              : launch_nn_t<BM_, BN_, WM_, WN_, true, false>(ctx, st, M, A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps)) \
        : (ab ? launch_nn_t<BM_, BN_, WM_, WN_, false, true>(ctx, st, M, A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps) \
              : launch_nn_t<BM_, BN_, WM_, WN_, false, false>(ctx, st, M, A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps)))
    return PB_NN(32, 64, 2, 4);
#undef PB_NN
}

template <bool BIG, bool AA, bool AB>
int launch_dw_t(pb_ctx* ctx, cudaStream_t st, int M, const double* A, int64_t lda, int64_t sA, int ka,
                const double* B, int64_t ldb, int64_t sB, int kb, int64_t rows, int64_t rps, int splits,
                double* out, int64_t ldo, int64_t split_stride, int64_t sOut, double* bias_out) {
    constexpr int BM = BIG ? 128 : 64, BN = BM;
    auto kern = train_dw_kernel<BM, BN, 2, 4, AA, AB>;
    constexpr size_t smem = (size_t)STAGES * BK * ((BM + 4) + (BN + 4)) * sizeof(double);
    if (!ctx) return set_smem(ctx, kern, smem);      // configuration pass (before any capture)
    const int tiles_m = (ka + BM - 1) / BM, tiles_n = (kb + BN - 1) / BN;
    dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)splits, (unsigned)M);
    PB_CUDA(ctx, launch_k(kern, grid, dim3(NTHREADS), smem, st, A, lda, sA, ka, B, ldb, sB, kb, rows, rps, out, ldo,
                          split_stride, sOut, bias_out, tiles_n));
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

int launch_dw(pb_ctx* ctx, cudaStream_t st, bool big, int M, const double* A, int64_t lda, int64_t sA, int ka,
              const double* B, int64_t ldb, int64_t sB, int kb, int64_t rows, int64_t rps, int splits,
              double* out, int64_t ldo, int64_t split_stride, int64_t sOut, double* bias_out) {
    const bool aa = al16(A, lda, sA), ab = al16(B, ldb, sB);
#define PB_DW(BIG_) \
    (aa ? (ab ? launch_dw_t<BIG_, true, true>(ctx, st, M, A, lda, sA, ka, B, ldb, sB, kb, rows, rps, splits, out, ldo, split_stride, sOut, bias_out) \
              : launch_dw_t<BIG_, true, false>(ctx, st, M, A, lda, sA, ka, B, ldb, sB, kb, rows, rps, splits, out, ldo, split_stride, sOut, bias_out)) \
        : (ab ? launch_dw_t<BIG_, false, true>(ctx, st, M, A, lda, sA, ka, B, ldb, sB, kb, rows, rps, splits, out, ldo, split_stride, sOut, bias_out) \
              : launch_dw_t<BIG_, false, false>(ctx, st, M, A, lda, sA, ka, B, ldb, sB, kb, rows, rps, splits, out, ldo, split_stride, sOut, bias_out)))
    if (big) return PB_DW(true);
    return PB_DW(false);
#undef PB_DW
}

// opt every instantiation into its dynamic shared memory once, outside any stream capture
template <int BM, int BN, int WM_, int WN_>
int configure_nn() {
    int rc = launch_nn_t<BM, BN, WM_, WN_, true, true>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.);
    if (rc == PB_OK) rc = launch_nn_t<BM, BN, WM_, WN_, true, false>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.);
    if (rc == PB_OK) rc = launch_nn_t<BM, BN, WM_, WN_, false, true>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.);
    if (rc == PB_OK) rc = launch_nn_t<BM, BN, WM_, WN_, false, false>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.);
    return rc;
}
int launch_fused(pb_ctx* ctx, struct TrainState& s);
int configure_fused();
int configure_all() {
    int rc = configure_nn<32, 64, 2, 4>();
    if (rc == PB_OK) rc = configure_fused();
    if (rc == PB_OK) rc = launch_dw_t<false, true, true>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    if (rc == PB_OK) rc = launch_dw_t<true, true, true>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    if (rc == PB_OK) rc = launch_dw_t<false, true, false>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    if (rc == PB_OK) rc = launch_dw_t<true, true, false>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    if (rc == PB_OK) rc = launch_dw_t<false, false, true>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    if (rc == PB_OK) rc = launch_dw_t<true, false, true>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    if (rc == PB_OK) rc = launch_dw_t<false, false, false>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    if (rc == PB_OK) rc = launch_dw_t<true, false, false>(nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    return rc;
}

int launch_adam(pb_ctx* ctx, TrainState& s, int update) {
    const TrainKey& k = s.key;
    dim3 grid((unsigned)((s.tab.P + 255) / 256), (unsigned)k.M);
    PB_CUDA(ctx, launch_k(train_adam_kernel, grid, dim3(256), 0, ctx->stream, s.tab, k.params, k.am, k.av,
                          s.WT, s.packed, (const double*)s.gpart, k.n_pass, k.use_adv ? 2 : 1, s.splits, k.M, k.lr, k.lam, s.tstep,
                          s.eidx, (const double*)s.nllpart, s.nll_blocks, s.regpart, s.ticket, s.loss, update));
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}


struct TrainState;
template <int MT>
int launch_fused_t(pb_ctx* ctx, TrainState* sp) {
    auto kern = train_fused_kernel<MT>;
    constexpr size_t smem = ((size_t)MT * FPITCH + (size_t)FRING * FCH) * sizeof(double) + 2 * FRING * sizeof(uint64_t) + 8 * sizeof(double);
    if (!ctx) return set_smem(ctx, kern, smem);
    TrainState& s = *sp; const TrainKey& k = s.key;
    FusedArgs a;
    a.M = k.M; a.L = k.n_layers; a.n_L = k.dims[k.n_layers] / 2; a.n_pass = k.n_pass; a.n_tiles = s.n_tiles; a.N = k.N;
    for (int l = 0; l <= k.n_layers; ++l) a.dims[l] = k.dims[l];
    a.packed = s.packed; a.SP = s.tab.SP;
    for (int l = 0; l < k.n_layers; ++l) { a.pk_f[l] = s.tab.pk_f[l]; a.pk_b[l] = s.tab.pk_b[l]; }
    a.X = k.X; a.v = k.v; a.xadv = s.xadv;
    for (int p = 0; p < 2; ++p) { a.act[p] = s.act[p]; a.del[p] = s.del[p]; }
    for (int l = 0; l <= k.n_layers; ++l) { a.act_off[l] = s.act_off[l]; a.del_off[l] = s.del_off[l]; }
    a.nllpart = s.nllpart; a.adv_eps = k.adv_eps; a.soft_0 = k.soft_0;
    PB_CUDA(ctx, launch_k(kern, dim3((unsigned)s.fused_grid), dim3(FUSED_THREADS), smem, ctx->stream, a));
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

int launch_fused(pb_ctx* ctx, TrainState& s) {
    return s.mt == 16 ? launch_fused_t<16>(ctx, &s) : launch_fused_t<32>(ctx, &s);
}

template <bool BIG>
int launch_dw_multi_t(pb_ctx* ctx, TrainState* sp) {
    constexpr int BM = BIG ? 128 : 64;
    auto kern = train_dw_multi_kernel<BM, BM, 2, 4>;
    constexpr size_t smem = (size_t)STAGES * BK * 2 * (BM + 4) * sizeof(double);
    if (!ctx) return set_smem(ctx, kern, smem);
    TrainState& s = *sp; const TrainKey& k = s.key;
    dim3 grid((unsigned)s.dw_tiles, (unsigned)s.splits, (unsigned)k.M);
    PB_CUDA(ctx, launch_k(kern, grid, dim3(NTHREADS), smem, ctx->stream, s.jobs, (int64_t)k.N, s.rps, s.gpart,
                          (int64_t)k.M * s.tab.P, (int64_t)s.tab.P));
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

int configure_fused() {
    int rc = launch_fused_t<16>(nullptr, nullptr);
    if (rc == PB_OK) rc = launch_fused_t<32>(nullptr, nullptr);
    if (rc == PB_OK) rc = launch_dw_multi_t<false>(nullptr, nullptr);
    if (rc == PB_OK) rc = launch_dw_multi_t<true>(nullptr, nullptr);
    return rc;
}

// fused epoch: 3 launches (forward + NLL + delta chain of both passes; all weight gradients; Adam)
int run_epoch_fused(pb_ctx* ctx, TrainState& s) {
    PB_TRY(launch_fused(ctx, s));
    PB_TRY(s.dw_big ? launch_dw_multi_t<true>(ctx, &s) : launch_dw_multi_t<false>(ctx, &s));
    PB_TRY(launch_adam(ctx, s, 1));
    return PB_OK;
}

// one epoch: main chain on ctx->stream, weight-gradient products on the side stream (forked per layer, joined
// before the Adam kernel).  Valid both eagerly and under stream capture.
int run_epoch(pb_ctx* ctx, TrainState& s) {
    const TrainKey& k = s.key;
    const int L = k.n_layers, M = k.M;
    const int64_t N = k.N;
    const int n_L = k.dims[L] / 2;
    cudaStream_t st = ctx->stream;
    const int n_pass = k.n_pass;
    for (int p = 0; p < n_pass; ++p) {
        const double* in0 = (p == 0) ? k.X : s.xadv;
        const int64_t s_in0 = (p == 0) ? 0 : N * k.dims[0];
        // forward
        for (int l = 0; l < L; ++l) {
            const double* A = (l == 0) ? in0 : s.act[p] + s.act_off[l];
            const int64_t sA = (l == 0) ? s_in0 : N * k.dims[l];
            const bool head = (l == L - 1);
            double* Y = head ? s.del[p] + s.del_off[l] : s.act[p] + s.act_off[l + 1];
            PB_TRY(launch_nn(ctx, st, M, A, k.dims[l], sA, N, k.dims[l], k.params + s.tab.off_w[l], k.dims[l + 1],
                             s.tab.P, k.dims[l + 1], Y, k.dims[l + 1], N * k.dims[l + 1],
                             head ? EPI_BIAS : EPI_BIAS_RELU, k.params + s.tab.off_b[l], 0, s.tab.P, 0.));
        }
        {   // NLL + head gradient, in place
            dim3 grid((unsigned)s.nll_blocks, (unsigned)M);
            PB_CUDA(ctx, launch_k(train_nll_kernel, grid, dim3(256), 0, st, s.del[p] + s.del_off[L - 1],
                                  (int64_t)(N * k.dims[L]), k.v, N, n_L, k.soft_0,
                                  s.nllpart + (int64_t)p * M * s.nll_blocks));
            PB_LAUNCH_CHECK(ctx);
        }
        // backward
        for (int l = L - 1; l >= 0; --l) {
            const double* D = s.del[p] + s.del_off[l];
            const int64_t sD = N * k.dims[l + 1];
            PB_CUDA(ctx, cudaEventRecord(s.ev_fork, st));
            PB_CUDA(ctx, cudaStreamWaitEvent(s.side, s.ev_fork, 0));
            {
                const double* A = (l == 0) ? in0 : s.act[p] + s.act_off[l];
                const int64_t sA = (l == 0) ? s_in0 : N * k.dims[l];
                double* gp = s.gpart + (int64_t)p * s.splits * M * s.tab.P;
                PB_TRY(launch_dw(ctx, s.side, s.dw_big, M, A, k.dims[l], sA, k.dims[l], D, k.dims[l + 1], sD, k.dims[l + 1], N,
                                 s.rps, s.splits, gp + s.tab.off_w[l], k.dims[l + 1], (int64_t)M * s.tab.P, s.tab.P,
                                 gp + s.tab.off_b[l]));
            }
            if (l > 0) {
                PB_TRY(launch_nn(ctx, st, M, D, k.dims[l + 1], sD, N, k.dims[l + 1], s.WT + s.tab.off_wt[l], k.dims[l],
                                 s.tab.PT, k.dims[l], s.del[p] + s.del_off[l - 1], k.dims[l], N * k.dims[l], EPI_MASK,
                                 s.act[p] + s.act_off[l], k.dims[l], N * k.dims[l], 0.));
            } else if (p == 0 && n_pass == 2) {
                PB_TRY(launch_nn(ctx, st, M, D, k.dims[1], sD, N, k.dims[1], s.WT + s.tab.off_wt[0], k.dims[0],
                                 s.tab.PT, k.dims[0], s.xadv, k.dims[0], N * k.dims[0], EPI_SIGN, k.X, k.dims[0], 0,
                                 k.adv_eps));
            }
        }
    }
    PB_CUDA(ctx, cudaEventRecord(s.ev_join, s.side));
    PB_CUDA(ctx, cudaStreamWaitEvent(st, s.ev_join, 0));
    PB_TRY(launch_adam(ctx, s, 1));
    return PB_OK;
}

int reserve(pb_ctx* ctx, double** p, size_t* cap, size_t elems) { return pb_reserve(ctx, p, cap, elems); }

}  // namespace

void pbi_train_free(pb_ctx* ctx) {
    TrainState* s = (TrainState*)ctx->train_state;
    if (!s) return;
    if (s->exec) cudaGraphExecDestroy(s->exec);
    for (int p = 0; p < 2; ++p) { cudaFree(s->act[p]); cudaFree(s->del[p]); }
    cudaFree(s->packed);
    cudaFree(s->xadv); cudaFree(s->WT); cudaFree(s->gpart); cudaFree(s->nllpart); cudaFree(s->regpart);
    cudaFree(s->loss); cudaFree(s->ticket); cudaFree(s->tstep); cudaFree(s->eidx);
    if (s->side) cudaStreamDestroy(s->side);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    delete s;
    ctx->train_state = nullptr;
}

extern "C" int pb_ensemble_train(pb_ctx* ctx, int n_members, int n_layers, const int* dims, double* params_dev,
                                 double* adam_m_dev, double* adam_v_dev, int64_t step0, int epochs, double lr,
                                 double lam, int use_adv, double adv_eps, double soft_0, const double* X_dev,
                                 const double* v_dev, int64_t N, double* loss_host, int use_graph) {
    if (!ctx) PB_FAIL(ctx, PB_ERR_ARG, "pb_ensemble_train: null context");
    if (n_members < 1 || n_layers < 1 || n_layers > MAXL || !dims || N < 1 || epochs < 0 || step0 < 0)
        PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_train: bad arguments (members %d, layers %d, N %lld, epochs %d)",
                n_members, n_layers, (long long)N, epochs);
    if (dims[n_layers] % 2 != 0) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_train: head width %d is not 2 n_L", dims[n_layers]);
    for (int l = 0; l <= n_layers; ++l)
        if (dims[l] < 1) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_train: layer width %d", dims[l]);
    if (!params_dev || !adam_m_dev || !adam_v_dev || !X_dev || !v_dev)
        PB_FAIL(ctx, PB_ERR_ARG, "pb_ensemble_train: null buffer");
    if (epochs == 0) return PB_OK;
    if (n_members > 64) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_train: at most 64 members per call (got %d)", n_members);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    {   // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: configure once per GPU, under a lock
        static std::mutex cfg_mu;
        static bool configured[64] = {false};
        std::lock_guard<std::mutex> lk(cfg_mu);
        const int d = ctx->device;
        if (d < 0 || d >= 64 || !configured[d]) {
            if (configure_all() != PB_OK) PB_FAIL(ctx, PB_ERR_CUDA, "pb_ensemble_train: kernel configuration failed: %s", pb_last_error(nullptr));
            if (d >= 0 && d < 64) configured[d] = true;
        }
    }

    if (!ctx->train_state) ctx->train_state = new TrainState();
    TrainState& s = *(TrainState*)ctx->train_state;
    TrainKey key; memset(&key, 0, sizeof key);
    key.M = n_members; key.n_layers = n_layers; key.use_adv = use_adv ? 1 : 0; key.N = N;
    for (int l = 0; l <= n_layers; ++l) key.dims[l] = dims[l];
    key.X = X_dev; key.v = v_dev; key.params = params_dev; key.am = adam_m_dev; key.av = adam_v_dev;
    key.n_pass = (use_adv && adv_eps != 0.) ? 2 : 1;
    key.lr = lr; key.lam = lam; key.adv_eps = use_adv ? adv_eps : 0.; key.soft_0 = soft_0;
    const bool same = s.valid && (s.key == key);
    const int M = n_members, L = n_layers;

    if (!same) {
        if (s.exec) { cudaGraphExecDestroy(s.exec); s.exec = nullptr; }
        memcpy(&s.key, &key, sizeof key); s.valid = false;
        LayerTab& t = s.tab; t.n_layers = L;
        long long off = 0, offt = 0;
        for (int l = 0; l < L; ++l) {
            t.in[l] = dims[l]; t.out[l] = dims[l + 1];
            t.off_w[l] = off; off += (long long)dims[l] * dims[l + 1];
            t.off_b[l] = off; off += dims[l + 1];
            t.off_wt[l] = offt; offt += (long long)dims[l] * dims[l + 1];
            offt += offt & 1;                                   // keep the transposed kernels 16-byte aligned
        }
        t.P = off; t.PT = offt + (offt & 1);
        // buffers
        s.act_total = 0; s.del_total = 0;
        // every segment starts 16-byte aligned (vector stores / 16-byte staging of even-width layers)
        for (int l = 1; l < L; ++l) { s.act_off[l] = s.act_total; s.act_total += (int64_t)M * N * dims[l]; s.act_total += s.act_total & 1; }
        for (int l = 0; l < L; ++l) { s.del_off[l] = s.del_total; s.del_total += (int64_t)M * N * dims[l + 1]; s.del_total += s.del_total & 1; }
        // split-K of the weight gradients: fill the SMs, at least 64 rows per split.  Large batches take 128 x 128
        // tiles (half the operand traffic per flop: the 64 x 64 tiles are HBM / L2 bound there).
        {
            static const int force = getenv("PB_TRAIN_DW") ? atoi(getenv("PB_TRAIN_DW")) : 0;   // 1 small, 2 big
            s.dw_big = force ? (force == 2) : ((int64_t)N * M >= 65536);
            const int bt = s.dw_big ? 128 : 64;
            int tiles = 0;
            for (int l = 0; l < L; ++l) tiles = std::max(tiles, ((dims[l] + bt - 1) / bt) * ((dims[l + 1] + bt - 1) / bt));
            int want = std::max(1, ctx->sm_count / std::max(1, tiles * M));
            want = std::min(want, MAX_SPLITS);
            int64_t rps = (N + want - 1) / want;
            rps = std::max<int64_t>(64, ((rps + BK - 1) / BK) * BK);
            s.rps = rps; s.splits = (int)((N + rps - 1) / rps);
        }
        const int n_pass = key.n_pass;
        {   // fused path: every width fits one 128-column chunk
            static const int want_fused = getenv("PB_TRAIN_FUSED") ? atoi(getenv("PB_TRAIN_FUSED")) : 1;
            bool ok = want_fused != 0;
            for (int l = 0; l <= L; ++l) ok = ok && dims[l] <= 128;
            s.fused = ok; t.fused = ok ? 1 : 0;
            long long offp = 0;
            for (int l = 0; l < L; ++l) {         // forward stream: K chunks of W_l, then its bias row
                t.pk_f[l] = offp; offp += (long long)((dims[l] + BK - 1) / BK) * FCH + FPITCH;
            }
            for (int l = L - 1; l >= 0; --l) {    // backward stream: K chunks of W_l^T
                t.pk_b[l] = offp; offp += (long long)((dims[l + 1] + BK - 1) / BK) * FCH;
            }
            t.SP = offp;
            if (s.fused) {
                // rows per tile: the largest of 64 / 32 / 16 that still gives every SM an item (weights are
                // re-streamed from L2 per item, so larger tiles halve that traffic)
                static const int force_mt = getenv("PB_TRAIN_MT") ? atoi(getenv("PB_TRAIN_MT")) : 0;
                // 32-row tiles (two co-resident CTAs per SM overlap each other's epilogues) once there are enough
                // items to fill them twice over, else 16-row tiles (measured: profiles/r01_notes.md)
                int mt = ((int64_t)M * ((N + 31) / 32) >= 4 * (int64_t)ctx->sm_count) ? 32 : 16;
                if (force_mt == 16 || force_mt == 32) mt = force_mt;
                s.mt = mt; s.n_tiles = (int)((N + mt - 1) / mt);
                static const int cps_env = getenv("PB_TRAIN_CPS") ? atoi(getenv("PB_TRAIN_CPS")) : 0;
                const int cps = cps_env ? cps_env : 2;                             // co-resident CTAs per SM
                s.fused_grid = (int)std::min<int64_t>((int64_t)M * s.n_tiles, (int64_t)ctx->sm_count * cps);
                s.nll_blocks = s.n_tiles;
                PB_TRY(reserve(ctx, &s.packed, &s.cap_packed, (size_t)M * t.SP));
                PB_CUDA(ctx, cudaMemsetAsync(s.packed, 0, (size_t)M * t.SP * sizeof(double), ctx->stream));
            }
        }
        for (int p = 0; p < n_pass; ++p) {
            PB_TRY(reserve(ctx, &s.act[p], &s.cap_act[p], (size_t)std::max<int64_t>(s.act_total, 2)));
            PB_TRY(reserve(ctx, &s.del[p], &s.cap_del[p], (size_t)s.del_total));
        }
        PB_TRY(reserve(ctx, &s.xadv, &s.cap_xadv, (size_t)M * N * dims[0] + 2));
        PB_TRY(reserve(ctx, &s.WT, &s.cap_WT, (size_t)M * t.PT));
        PB_TRY(reserve(ctx, &s.gpart, &s.cap_gpart, (size_t)n_pass * s.splits * M * t.P));
        if (!s.fused) s.nll_blocks = (int)std::min<int64_t>(NLL_BLOCKS_MAX, std::max<int64_t>(1, (N * (dims[L] / 2) + 511) / 512));
        PB_TRY(reserve(ctx, &s.nllpart, &s.cap_nll, (size_t)2 * M * std::max(NLL_BLOCKS_MAX, s.nll_blocks)));
        if (s.fused) {   // job table of the batched weight-gradient launch: (pass, layer)
            const int bt = s.dw_big ? 128 : 64;
            s.jobs.n_jobs = 0; int tile0 = 0;
            for (int p = 0; p < n_pass; ++p)
                for (int l = 0; l < L; ++l) {
                    DwJob& J = s.jobs.job[s.jobs.n_jobs++];
                    J.A = (l == 0) ? (p == 0 ? X_dev : s.xadv) : s.act[p] + s.act_off[l];
                    J.sA = (l == 0) ? (p == 0 ? 0 : N * dims[0]) : N * dims[l];
                    J.lda = dims[l]; J.ka = dims[l];
                    J.B = s.del[p] + s.del_off[l]; J.sB = N * dims[l + 1]; J.ldb = dims[l + 1]; J.kb = dims[l + 1];
                    J.out_off = (long long)p * s.splits * M * t.P + t.off_w[l];
                    J.bias_off = (long long)p * s.splits * M * t.P + t.off_b[l];
                    J.tiles_n = (dims[l + 1] + bt - 1) / bt;
                    J.tile0 = tile0; tile0 += ((dims[l] + bt - 1) / bt) * J.tiles_n;
                    J.alignA = al16(J.A, J.lda, J.sA) ? 1 : 0; J.alignB = al16(J.B, J.ldb, J.sB) ? 1 : 0;
                }
            s.dw_tiles = tile0;
        }
        PB_TRY(reserve(ctx, &s.regpart, &s.cap_reg, (size_t)M * ((t.P + 255) / 256)));
        if (!s.ticket) {
            PB_CUDA(ctx, cudaMalloc(&s.ticket, 64 * sizeof(unsigned)));
            PB_CUDA(ctx, cudaMalloc(&s.tstep, 64 * sizeof(long long)));
            PB_CUDA(ctx, cudaMalloc(&s.eidx, 64 * sizeof(int)));
            PB_CUDA(ctx, cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
            PB_CUDA(ctx, cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
            PB_CUDA(ctx, cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming));
        }
        s.launches_per_epoch = s.fused ? 3 : n_pass * (L + 1 + L + (L - 1)) + (n_pass == 2 ? 1 : 0) + 1;
        s.valid = true;
    }
    // the loss history lives in a context buffer (its address is baked into the captured graph)
    if ((size_t)epochs * M > s.cap_loss) {
        if (s.exec) { cudaGraphExecDestroy(s.exec); s.exec = nullptr; }
        PB_TRY(reserve(ctx, &s.loss, &s.cap_loss, (size_t)epochs * M));
    }
    cudaStream_t st = ctx->stream;
    {   // per-call counters: Adam step (bias correction) and loss slot
        std::vector<long long> t0(64, step0);
        PB_CUDA(ctx, cudaMemcpyAsync(s.tstep, t0.data(), 64 * sizeof(long long), cudaMemcpyHostToDevice, st));
        PB_CUDA(ctx, cudaMemsetAsync(s.eidx, 0, 64 * sizeof(int), st));
        PB_CUDA(ctx, cudaMemsetAsync(s.ticket, 0, 64 * sizeof(unsigned), st));
        PB_CUDA(ctx, cudaStreamSynchronize(st));      // t0 is a host temporary
    }
    PB_TRY(launch_adam(ctx, s, 0));                   // transposed kernels of the incoming parameters

    if (use_graph) {
        if (!s.exec) {
            cudaGraph_t graph = nullptr;
            const int64_t before = ctx->launches;
            PB_CUDA(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
            int rc = s.fused ? run_epoch_fused(ctx, s) : run_epoch(ctx, s);
            cudaError_t e = cudaStreamEndCapture(st, &graph);
            ctx->launches = before;
            if (rc != PB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) PB_FAIL(ctx, PB_ERR_CUDA, "pb_ensemble_train: graph capture -> %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&s.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { s.exec = nullptr; PB_FAIL(ctx, PB_ERR_CUDA, "pb_ensemble_train: graph instantiate -> %s", cudaGetErrorString(e)); }
        }
        for (int ep = 0; ep < epochs; ++ep) PB_CUDA(ctx, cudaGraphLaunch(s.exec, st));
        ctx->launches += (int64_t)epochs * s.launches_per_epoch;
    } else {
        for (int ep = 0; ep < epochs; ++ep) PB_TRY(s.fused ? run_epoch_fused(ctx, s) : run_epoch(ctx, s));
    }
    if (loss_host) {
        PB_CUDA(ctx, cudaMemcpyAsync(loss_host, s.loss, (size_t)epochs * M * sizeof(double), cudaMemcpyDeviceToHost, st));
        PB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return PB_OK;
}
