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
};

enum { EPI_BIAS_RELU = 0, EPI_BIAS = 1, EPI_MASK = 2, EPI_SIGN = 3 };

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
                  double* __restrict__ WT, const double* __restrict__ gpart, int n_pass, int n_term, int splits, int M,
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
            // transposed copy of the kernels
            for (int l = 0; l < tab.n_layers; ++l) {
                const long long q = p - tab.off_w[l];
                if (q >= 0 && q < (long long)tab.in[l] * tab.out[l]) {
                    const int i = (int)(q / tab.out[l]), j = (int)(q % tab.out[l]);
                    WT[(long long)mem * tab.PT + tab.off_wt[l] + (long long)j * tab.in[l] + i] = wn;
                    break;
                }
            }
        } else {
            for (int l = 0; l < tab.n_layers; ++l) {
                const long long q = p - tab.off_w[l];
                if (q >= 0 && q < (long long)tab.in[l] * tab.out[l]) {
                    const int i = (int)(q / tab.out[l]), j = (int)(q % tab.out[l]);
                    WT[(long long)mem * tab.PT + tab.off_wt[l] + (long long)j * tab.in[l] + i] = w0;
                    break;
                }
            }
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
    PB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
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
    kern<<<grid, NTHREADS, smem, st>>>(A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

int launch_nn(pb_ctx* ctx, cudaStream_t st, int M, const double* A, int64_t lda, int64_t sA, int64_t rows, int m,
              const double* B, int64_t ldb, int64_t sB, int k, double* Y, int64_t ldy, int64_t sY, int epi,
              const double* aux, int64_t ldaux, int64_t sAux, double eps) {
    const bool aa = al16(A, lda, sA), ab = al16(B, ldb, sB);
    const int64_t big_ctas = ((rows + 127) / 128) * ((k + 63) / 64) * M;
    static const int force = getenv("PB_TRAIN_TILE") ? atoi(getenv("PB_TRAIN_TILE")) : 0;   // 1 small, 2 big
    (void)big_ctas;
    const bool big = (force == 2);     // measured: the 32-row tiles (3 CTAs per SM) win at every N (profiles/r01_notes.md)
#define PB_NN(BM_, BN_, WM_, WN_) \
    (aa ? (ab ? launch_nn_t<BM_, BN_, WM_, WN_, true, true>(ctx, st, M, A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps) \
              : launch_nn_t<BM_, BN_, WM_, WN_, true, false>(ctx, st, M, A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps)) \
        : (ab ? launch_nn_t<BM_, BN_, WM_, WN_, false, true>(ctx, st, M, A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps) \
              : launch_nn_t<BM_, BN_, WM_, WN_, false, false>(ctx, st, M, A, lda, sA, rows, m, B, ldb, sB, k, Y, ldy, sY, epi, aux, ldaux, sAux, eps)))
    if (big) return PB_NN(128, 64, 4, 2);
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
    kern<<<grid, NTHREADS, smem, st>>>(A, lda, sA, ka, B, ldb, sB, kb, rows, rps, out, ldo, split_stride, sOut,
                                       bias_out, tiles_n);
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
int configure_all() {
    int rc = configure_nn<128, 64, 4, 2>();
    if (rc == PB_OK) rc = configure_nn<32, 64, 2, 4>();
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
    train_adam_kernel<<<grid, 256, 0, ctx->stream>>>(s.tab, k.params, k.am, k.av, s.WT, s.gpart, k.n_pass,
                                                     k.use_adv ? 2 : 1, s.splits, k.M, k.lr, k.lam, s.tstep, s.eidx, s.nllpart,
                                                     s.nll_blocks, s.regpart, s.ticket, s.loss, update);
    PB_LAUNCH_CHECK(ctx);
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
            train_nll_kernel<<<grid, 256, 0, st>>>(s.del[p] + s.del_off[L - 1], N * k.dims[L], k.v, N, n_L, k.soft_0,
                                                   s.nllpart + (int64_t)p * M * s.nll_blocks);
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
    {
        static bool configured = false;
        if (!configured) {
            if (configure_all() != PB_OK) PB_FAIL(ctx, PB_ERR_CUDA, "pb_ensemble_train: kernel configuration failed: %s", pb_last_error(nullptr));
            configured = true;
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
        for (int l = 1; l < L; ++l) { s.act_off[l] = s.act_total; s.act_total += (int64_t)M * N * dims[l]; }
        for (int l = 0; l < L; ++l) { s.del_off[l] = s.del_total; s.del_total += (int64_t)M * N * dims[l + 1]; }
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
        for (int p = 0; p < n_pass; ++p) {
            PB_TRY(reserve(ctx, &s.act[p], &s.cap_act[p], (size_t)std::max<int64_t>(s.act_total, 2)));
            PB_TRY(reserve(ctx, &s.del[p], &s.cap_del[p], (size_t)s.del_total));
        }
        PB_TRY(reserve(ctx, &s.xadv, &s.cap_xadv, (size_t)M * N * dims[0] + 2));
        PB_TRY(reserve(ctx, &s.WT, &s.cap_WT, (size_t)M * t.PT));
        PB_TRY(reserve(ctx, &s.gpart, &s.cap_gpart, (size_t)n_pass * s.splits * M * t.P));
        s.nll_blocks = (int)std::min<int64_t>(NLL_BLOCKS_MAX, std::max<int64_t>(1, (N * (dims[L] / 2) + 511) / 512));
        PB_TRY(reserve(ctx, &s.nllpart, &s.cap_nll, (size_t)2 * M * NLL_BLOCKS_MAX));
        PB_TRY(reserve(ctx, &s.regpart, &s.cap_reg, (size_t)M * ((t.P + 255) / 256)));
        if (!s.ticket) {
            PB_CUDA(ctx, cudaMalloc(&s.ticket, 64 * sizeof(unsigned)));
            PB_CUDA(ctx, cudaMalloc(&s.tstep, 64 * sizeof(long long)));
            PB_CUDA(ctx, cudaMalloc(&s.eidx, 64 * sizeof(int)));
            PB_CUDA(ctx, cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
            PB_CUDA(ctx, cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
            PB_CUDA(ctx, cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming));
        }
        s.launches_per_epoch = n_pass * (L + 1 + L + (L - 1)) + (n_pass == 2 ? 1 : 0) + 1;
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
            int rc = run_epoch(ctx, s);
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
        for (int ep = 0; ep < epochs; ++ep) PB_TRY(run_epoch(ctx, s));
    }
    if (loss_host) {
        PB_CUDA(ctx, cudaMemcpyAsync(loss_host, s.loss, (size_t)epochs * M * sizeof(double), cudaMemcpyDeviceToHost, st));
        PB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return PB_OK;
}
