// Deep-ensemble surrogate forward: M mean/variance MLPs evaluated on a batch of parameter
// points, fused with the Gaussian-mixture moments of PodnnModel.predict_v.
//
// Replaces varneuralnetwork.py:42-62 (Dense+ReLU stack, linear 2*n_L head, softplus scale),
// :75-86 (input normalisation), :154-160 (mean / variance) and podnnmodel.py:314-328
// (mean over members of mu and of sigma^2 + mu^2) -- serial over members and layers in the
// reference, with (N, n_L, M) intermediates.
//
// One persistent CTA owns a tile of 128 points.  Activations stay in shared memory
// ([128][132] doubles, in place across layers); the layer weights stream through a 4-stage
// cp.async ring in 16-row chunks (they are L2 resident: ~0.4 MB per member); the layer
// product runs on DMMA with the accumulators in registers; bias + ReLU are applied in the
// epilogue that writes the next layer's activations.  The head epilogue adds mu, mu^2 and
// sigma^2 of the member into three running sums; after the last member the tile is
// finalised to (mean, sqrt(var)).  No (N, n_L, M) intermediate ever reaches HBM.
#include "common.cuh"
#include "tile_ops.cuh"

namespace {
using namespace pbt;

constexpr int MP = 128;            // points per tile
constexpr int PACT = 132;          // activation pitch (== 4 mod 16: conflict-free A fragments)
constexpr int PW = 132;            // weight-stage pitch
constexpr int STAGES = 4;
constexpr int MAX_LAYERS = 8;
constexpr int MAX_WIDTH = 128;     // hidden widths (inputs of every layer) must fit the tile

struct MlpArgs {
    int n_members, n_layers, n_L, norm_kind;
    int dims[MAX_LAYERS + 1];
    int64_t w_off[MAX_LAYERS], b_off[MAX_LAYERS], member_stride;
    const double* weights; const double* norm_a; const double* norm_b;
    double soft0;
    const double* X; int64_t N;
    double* s1; double* s2; double* s3;     // running sums: mu, mu^2, sigma^2  (N x n_L each)
    int64_t n_tiles;
};

__device__ __forceinline__ double softplus(double x) {
    // np.logaddexp(0, x) == tf.math.softplus within 1 ulp
    return x > 0. ? x + log1p(exp(-x)) : log1p(exp(x));
}

__global__ void __launch_bounds__(NTHREADS, 1) ensemble_kernel(MlpArgs a) {
    extern __shared__ __align__(16) double smem[];
    double* act = smem;                          // [MP][PACT]
    double* Ws = smem + MP * PACT;               // [STAGES][BK][PW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;     // 2 x 4 warps, warp tile 64 x 32
    const int lk = lane & 3, lg = lane >> 2;
    const int n_d = a.dims[0], n_L = a.n_L;

    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * MP;
        for (int mem = 0; mem < a.n_members; ++mem) {
            const double* wbase = a.weights + (int64_t)mem * a.member_stride;
            // ---- input tile, normalised (varneuralnetwork.py:75-86), zero padded to 16 columns
            __syncthreads();
            const int kin0 = (n_d + BK - 1) / BK * BK;
            for (int e = tid; e < MP * kin0; e += NTHREADS) {
                int p = e / kin0, c = e % kin0;
                double v = 0.;
                if (c < n_d && p0 + p < a.N) {
                    v = a.X[(p0 + p) * n_d + c];
                    if (a.norm_kind == PB_NORM_MEANSTD) v = (v - a.norm_a[c]) / a.norm_b[c];
                    else if (a.norm_kind == PB_NORM_CENTER) v = (v - a.norm_a[c]) - 0.5 * (a.norm_b[c] - a.norm_a[c]);
                }
                act[p * PACT + c] = v;
            }
            __syncthreads();
            for (int l = 0; l < a.n_layers; ++l) {
                const int in = a.dims[l], out = a.dims[l + 1];
                const bool head = (l == a.n_layers - 1);
                const double* W = wbase + a.w_off[l];
                const double* bias = wbase + a.b_off[l];
                const bool al = ((reinterpret_cast<uintptr_t>(W) % 16) == 0) && (out % 2 == 0);
                const int nk = (in + BK - 1) / BK;
                for (int c0 = 0; c0 < out; c0 += 128) {       // > 1 chunk only for a wide head
                    double acc[8][4][2];
#pragma unroll
                    for (int t = 0; t < 8; ++t)
#pragma unroll
                        for (int u = 0; u < 4; ++u) { acc[t][u][0] = 0.; acc[t][u][1] = 0.; }
#pragma unroll
                    for (int s = 0; s < STAGES - 1; ++s) {
                        if (s < nk) {
                            if (al) load_tile<BK, 128, PW, true>(Ws + s * BK * PW, W, out, (int64_t)s * BK, in, c0, out, tid);
                            else    load_tile<BK, 128, PW, false>(Ws + s * BK * PW, W, out, (int64_t)s * BK, in, c0, out, tid);
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
                                if (al) load_tile<BK, 128, PW, true>(Ws + s * BK * PW, W, out, (int64_t)kn * BK, in, c0, out, tid);
                                else    load_tile<BK, 128, PW, false>(Ws + s * BK * PW, W, out, (int64_t)kn * BK, in, c0, out, tid);
                            }
                            cp_async_commit();
                        }
                        const double* as = act + (wm * 64 + lg) * PACT + kt * BK + lk;
                        const double* bs = Ws + (kt % STAGES) * BK * PW + wn * 32 + lg;
#pragma unroll
                        for (int kk = 0; kk < BK / 4; ++kk) {
                            double af[8], bf[4];
#pragma unroll
                            for (int t = 0; t < 8; ++t) af[t] = as[t * 8 * PACT + kk * 4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) bf[u] = bs[(kk * 4 + lk) * PW + u * 8];
#pragma unroll
                            for (int t = 0; t < 8; ++t)
#pragma unroll
                                for (int u = 0; u < 4; ++u) dmma884(acc[t][u][0], acc[t][u][1], af[t], bf[u]);
                        }
                    }
                    cp_async_wait<0>();
                    __syncthreads();            // every warp is done reading act and the weight ring
                    if (!head) {
                        // next activations: relu(acc + b), columns [out, ceil16(out)) zeroed
                        const int outp = (out + BK - 1) / BK * BK;
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int p = wm * 64 + t * 8 + lg;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    const int n = wn * 32 + u * 8 + 2 * lk + q;
                                    if (n < out) act[p * PACT + n] = fmax(acc[t][u][q] + bias[n], 0.);
                                    else if (n < outp) act[p * PACT + n] = 0.;
                                }
                            }
                        }
                    } else {
                        // head: loc = t[:, :n_L]; scale = softplus(soft_0 * t[:, n_L:]) + 1e-6
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int64_t p = p0 + wm * 64 + t * 8 + lg;
                            if (p >= a.N) continue;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    const int n = c0 + wn * 32 + u * 8 + 2 * lk + q;
                                    if (n >= out) continue;
                                    const double v = acc[t][u][q] + bias[n];
                                    if (n < n_L) {
                                        const int64_t o = p * n_L + n;
                                        if (mem == 0) { a.s1[o] = v; a.s2[o] = v * v; }
                                        else { a.s1[o] += v; a.s2[o] += v * v; }
                                    } else {
                                        const int64_t o = p * n_L + (n - n_L);
                                        const double sc = softplus(a.soft0 * v) + 1e-6;
                                        if (mem == 0) a.s3[o] = sc * sc; else a.s3[o] += sc * sc;
                                    }
                                }
                            }
                        }
                    }
                    __syncthreads();
                }
            }
        }
        // ---- mixture moments (podnnmodel.py:322-326): mean, sqrt(mean(var + mu^2) - mean^2)
        __syncthreads();
        const double invM = 1. / (double)a.n_members;
        const int64_t pts = min((int64_t)MP, a.N - p0);
        for (int64_t e = tid; e < pts * n_L; e += NTHREADS) {
            const int64_t o = p0 * n_L + e;
            const double mean = a.s1[o] * invM;
            // one member: its own variance, exactly (VarNeuralNetwork.predict, varneuralnetwork.py:159)
            const double var = a.n_members == 1 ? a.s3[o] : (a.s3[o] + a.s2[o]) * invM - mean * mean;
            a.s1[o] = mean;
            a.s2[o] = sqrt(var);
        }
    }
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

}  // namespace

extern "C" int pb_ensemble_predict(pb_ctx* ctx, int n_members, int n_layers, const int* dims,
                                   const double* weights_dev, int norm_kind, const double* norm_a_dev,
                                   const double* norm_b_dev, double soft_0, const double* X_dev, int64_t N,
                                   double* v_mean_dev, double* v_sig_dev) {
    if (!ctx || !dims) return PB_ERR_ARG;
    if (n_members <= 0 || n_layers < 1 || n_layers > MAX_LAYERS || N <= 0)
        PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_predict: need 1..%d layers, >= 1 member, N > 0", MAX_LAYERS);
    const int out_last = dims[n_layers];
    if (out_last % 2) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_predict: the head must have 2*n_L outputs");
    for (int l = 0; l < n_layers; ++l)
        if (dims[l] <= 0 || dims[l] > MAX_WIDTH)
            PB_FAIL(ctx, PB_ERR_SHAPE, "pb_ensemble_predict: layer input width %d outside (0, %d]", dims[l], MAX_WIDTH);
    if (norm_kind != PB_NORM_NONE && (!norm_a_dev || !norm_b_dev)) PB_FAIL(ctx, PB_ERR_ARG, "pb_ensemble_predict: normalisation bounds missing");
    MlpArgs a{};
    a.n_members = n_members; a.n_layers = n_layers; a.n_L = out_last / 2; a.norm_kind = norm_kind;
    int64_t off = 0;
    for (int l = 0; l <= n_layers; ++l) a.dims[l] = dims[l];
    for (int l = 0; l < n_layers; ++l) {
        a.w_off[l] = off; off += (int64_t)dims[l] * dims[l + 1];
        a.b_off[l] = off; off += dims[l + 1];
    }
    a.member_stride = off;
    a.weights = weights_dev; a.norm_a = norm_a_dev; a.norm_b = norm_b_dev; a.soft0 = soft_0;
    a.X = X_dev; a.N = N; a.s1 = v_mean_dev; a.s2 = v_sig_dev;
    PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)N * a.n_L));
    a.s3 = ctx->ws;
    a.n_tiles = (N + MP - 1) / MP;
    cudaEventRecord(ctx->ev0, ctx->stream);
    size_t smem = sizeof(double) * (size_t)(MP * PACT + STAGES * BK * PW);
    PB_CUDA(ctx, cudaFuncSetAttribute(ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (int)std::min<int64_t>(a.n_tiles, ctx->sm_count);
    ensemble_kernel<<<grid, NTHREADS, smem, ctx->stream>>>(a);
    PB_LAUNCH_CHECK(ctx);
    cudaEventRecord(ctx->ev1, ctx->stream);
    PB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->ms[PB_T_PREDICT] = ms;
    return PB_OK;
}

extern "C" int pb_predict_U(pb_ctx* ctx, const double* V_dev, int64_t ldv, int64_t n_rows, int n_L,
                            const double* v_mean_dev, const double* v_sig_dev, int64_t N,
                            int samples, uint64_t seed, double* U_mean_dev, double* U_sig_dev, int64_t ldo) {
    if (!ctx) return PB_ERR_ARG;
    (void)seed;
    if (samples != 0) PB_FAIL(ctx, PB_ERR_ARG, "pb_predict_U: only the closed-form moments (samples = 0) are built in this round");
    if (n_L <= 0 || N <= 0 || n_rows <= 0 || ldo < N) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_predict_U: bad shape");
    // U_mean = V (n_rows x n_L) * v_mean^T (n_L x N)
    const int64_t ldb = (N + 1) / 2 * 2;
    PB_TRY(pb_reserve(ctx, &ctx->Bmat, &ctx->B_cap, (size_t)n_L * ldb));
    if (U_mean_dev) {
        PB_TRY(pbi_transpose(ctx, v_mean_dev, N, n_L, n_L, ctx->Bmat, ldb));
        PB_TRY(pbi_gemm_nn(ctx, V_dev, ldv, n_rows, n_L, ctx->Bmat, ldb, N, U_mean_dev, ldo, nullptr, 0, nullptr));
    }
    if (U_sig_dev) {
        // (V*V) (sig^2)^T, then sqrt
        double* V2 = nullptr; double* s2 = nullptr;
        PB_CUDA(ctx, cudaMalloc(&V2, sizeof(double) * (size_t)n_rows * n_L));
        PB_CUDA(ctx, cudaMalloc(&s2, sizeof(double) * (size_t)N * n_L));
        int64_t nv = n_rows * n_L, ns = N * n_L;
        square_strided_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, ctx->stream>>>(V_dev, ldv, V2, n_rows, n_L);
        PB_LAUNCH_CHECK(ctx);
        square_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, ctx->stream>>>(v_sig_dev, s2, ns);
        PB_LAUNCH_CHECK(ctx);
        int rc = pbi_transpose(ctx, s2, N, n_L, n_L, ctx->Bmat, ldb);
        if (rc == PB_OK) rc = pbi_gemm_nn(ctx, V2, n_L, n_rows, n_L, ctx->Bmat, ldb, N, U_sig_dev, ldo, nullptr, 0, nullptr);
        if (rc == PB_OK) {
            int64_t n = n_rows * N;
            sqrt_inplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(U_sig_dev, n_rows, N, ldo);
            ctx->launches++;
        }
        cudaStreamSynchronize(ctx->stream);
        cudaFree(V2); cudaFree(s2);
        if (rc != PB_OK) return rc;
    }
    return PB_OK;
}
