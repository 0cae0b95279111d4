// Measurement helpers used by bench.py: live fp64 peak (DFMA / DMMA / both pipes), HBM copy
// bandwidth and an L2 flush.  MEASURED_PEAKS.json has no fp64 entry, so the roofline
// denominator of the Gram kernel is measured in the same process as the kernel itself.
#include "common.cuh"
#include "tile_ops.cuh"

namespace {

using pbt::dmma884;

constexpr int PROBE_ITERS = 2048;

// kind 0: 16 independent DFMA chains per thread; kind 1: 16 independent DMMA accumulators per warp;
// kind 2: even warps DMMA, odd warps DFMA (are the two pipes separate?)
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int kind, double seed) {
    const int warp = threadIdx.x >> 5;
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-9 * (threadIdx.x & 7);
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i][0] = i * 1e-3; acc[i][1] = -i * 1e-3; }
    const bool use_mma = (kind == 1) || (kind == 2 && (warp & 1) == 0);
    if (use_mma) {
        for (int it = 0; it < PROBE_ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
        }
    } else {
        for (int it = 0; it < PROBE_ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                acc[i][0] = fma(acc[i][0], b, a);
                acc[i][1] = fma(acc[i][1], b, a);
            }
        }
    }
    double s = 0.;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;      // keep the work alive
}

__global__ void copy_kernel(const double4* __restrict__ in, double4* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}
__global__ void fill_kernel(double4* __restrict__ out, size_t n, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = make_double4(v, v, v, v);
}

}  // namespace

extern "C" int pb_probe_fp64_peak(pb_ctx* ctx, int kind, double* tflops) {
    PB_ENTER(ctx); if (!tflops || kind < 0 || kind > 2) return PB_ERR_ARG;
    double* out = nullptr;
    PB_CUDA(ctx, cudaMalloc(&out, 64));
    const int blocks = ctx->sm_count * 4;
    double best = 0.;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        fp64_probe_kernel<<<blocks, 256, 0, ctx->stream>>>(out, kind, 0.5);
        PB_LAUNCH_CHECK(ctx);
        cudaEventRecord(ctx->ev1, ctx->stream);
        PB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        // flops per warp per iteration: DMMA 16 x (8*8*4*2) = 8192; DFMA 32 lanes x 32 fma x 2 = 2048
        double warps = (double)blocks * 8.;
        double fl;
        if (kind == 0) fl = warps * PROBE_ITERS * 2048.;
        else if (kind == 1) fl = warps * PROBE_ITERS * 8192.;
        else fl = warps * 0.5 * PROBE_ITERS * (2048. + 8192.);
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(out);
    *tflops = best;
    return PB_OK;
}

extern "C" int pb_probe_hbm_copy(pb_ctx* ctx, size_t bytes, double* gbs) {
    PB_ENTER(ctx); if (!gbs) return PB_ERR_ARG;
    bytes = bytes / 32 * 32;
    void *a = nullptr, *b = nullptr;
    PB_CUDA(ctx, cudaMalloc(&a, bytes)); PB_CUDA(ctx, cudaMalloc(&b, bytes));
    size_t n = bytes / 32;
    fill_kernel<<<ctx->sm_count * 8, 512, 0, ctx->stream>>>((double4*)a, n, 1.0);
    PB_LAUNCH_CHECK(ctx);
    double best = 0.;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        copy_kernel<<<ctx->sm_count * 8, 512, 0, ctx->stream>>>((const double4*)a, (double4*)b, n);
        PB_LAUNCH_CHECK(ctx);
        cudaEventRecord(ctx->ev1, ctx->stream);
        PB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        double g = 2. * (double)bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaFree(a); cudaFree(b);
    *gbs = best;
    return PB_OK;
}

extern "C" int pb_flush_l2(pb_ctx* ctx) {
    PB_ENTER(ctx);
    const size_t bytes = (size_t)256 << 20;     // 2 x the 126 MB L2
    if (!ctx->flush) { PB_CUDA(ctx, cudaMalloc(&ctx->flush, bytes)); ctx->flush_bytes = bytes; }
    fill_kernel<<<ctx->sm_count * 8, 512, 0, ctx->stream>>>((double4*)ctx->flush, bytes / 32, 0.0);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}
