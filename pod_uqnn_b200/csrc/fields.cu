// Snapshot evaluation: the analytic fields u(X, t, mu) written straight into the
// row-major snapshot matrix U (n_h x n_st) in HBM.
//
// Replaces the per-sample Python/numba loops of poduqnn/acceleration.py:13-30 (loop_u)
// and :34-70 (loop_u_t), noise-free branch, for the registered fields.  One thread owns
// one snapshot COLUMN (sample i, time step j) and walks a chunk of mesh nodes, so a warp
// writes 32 consecutive doubles of each row (coalesced 256-byte stores); the reference
// writes strided columns.  Arithmetic follows the reference expression trees operation by
// operation (explicit _rn intrinsics keep ptxas from contracting a*b+c into FMA), so the
// only differences to numpy are the <= 1-2 ulp of exp / cos.
#include "common.cuh"

namespace {

constexpr int COLS_PER_BLOCK = 128;
constexpr int NODES_PER_BLOCK = 64;

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }

// Shekel, experiments/1d_shekel/hyperparams.py:55-67: sum_k 1 / (bet_k + (x - gam_k)^2)
__global__ void shekel_kernel(const double* __restrict__ X, int64_t n_rows, int64_t row0,
                              const double* __restrict__ mu, int64_t n_s, int n_mu,
                              double* __restrict__ U, int64_t ldu, int npb) {
    extern __shared__ double smu[];           // [n_mu][COLS_PER_BLOCK]
    const int64_t c0 = (int64_t)blockIdx.y * COLS_PER_BLOCK;
    for (int e = threadIdx.x; e < n_mu * COLS_PER_BLOCK; e += blockDim.x) {
        int c = e / n_mu, k = e % n_mu;
        smu[k * COLS_PER_BLOCK + c] = (c0 + c < n_s) ? mu[(c0 + c) * n_mu + k] : 1.;
    }
    __syncthreads();
    const int64_t col = c0 + threadIdx.x;
    const int half = n_mu / 2;
    const int64_t r_begin = (int64_t)blockIdx.x * npb;
    const int64_t r_end = min(n_rows, r_begin + npb);
    if (col >= n_s) return;
    for (int64_t r = r_begin; r < r_end; ++r) {
        const double x = X[row0 + r];
        double s = 0.;
        for (int k = 0; k < half; ++k) {
            double d = sub(x, smu[(half + k) * COLS_PER_BLOCK + threadIdx.x]);
            double q = mul(d, d);
            s = add(s, dvd(1., add(smu[k * COLS_PER_BLOCK + threadIdx.x], q)));
        }
        U[r * ldu + col] = s;
    }
}

// Ackley, experiments/2d_ackley/hyperparams.py:51-57
__global__ void ackley_kernel(const double* __restrict__ X, int64_t n_xyz, int64_t n_rows, int64_t row0,
                              const double* __restrict__ mu, int64_t n_s, int n_mu,
                              double* __restrict__ U, int64_t ldu) {
    const int64_t col = (int64_t)blockIdx.y * COLS_PER_BLOCK + threadIdx.x;
    const int64_t r_begin = (int64_t)blockIdx.x * NODES_PER_BLOCK;
    const int64_t r_end = min(n_rows, r_begin + NODES_PER_BLOCK);
    __shared__ double sx[NODES_PER_BLOCK], sy[NODES_PER_BLOCK], sr[NODES_PER_BLOCK];
    for (int e = threadIdx.x; e < NODES_PER_BLOCK; e += blockDim.x) {
        int64_t r = r_begin + e;
        if (r < r_end) {
            double x = X[row0 + r], y = X[n_xyz + row0 + r];
            sx[e] = x; sy[e] = y;
            sr[e] = sqrt(mul(.5, add(mul(x, x), mul(y, y))));      // np.sqrt(.5*(x**2+y**2))
        }
    }
    __syncthreads();
    if (col >= n_s) return;
    const double m0 = mu[col * n_mu + 0], m1 = mu[col * n_mu + 1], m2 = mu[col * n_mu + 2];
    const double a2 = mul(-20., add(1., mul(.1, m2)));             // -20*(1+.1*mu[2])
    const double c1 = mul(-.2, add(1., mul(.1, m1)));              // -.2*(1+.1*mu[1])
    const double w = mul(6.283185307179586, add(1., mul(.1, m0))); // 2*np.pi*(1+.1*mu[0])
    const double e1 = 2.718281828459045;                           // np.exp(1)
    for (int64_t r = r_begin; r < r_end; ++r) {
        const int e = (int)(r - r_begin);
        double t1 = mul(a2, exp(mul(c1, sr[e])));
        double t2 = exp(mul(.5, add(cos(mul(w, sx[e])), cos(mul(w, sy[e])))));
        U[r * ldu + col] = add(add(sub(t1, t2), 20.), e1);
    }
}

// Ackley with de-duplicated coordinates.  The second term exp(.5 (cos(w x) + cos(w y))) only depends on
// (x, sample) and (y, sample); on a tensor grid there are n_x + n_y distinct coordinates for n_x n_y nodes,
// so E[u][s] = exp(.5 cos(w_s u)) is tabulated once (L2 resident) and the term becomes a product of two
// table entries: one exp per element instead of two exp + two cos, which moves the kernel from the fp64
// issue limit to the HBM write roofline.  exp(a) exp(b) vs exp(a + b): <= 3 ulp of a term <= e, i.e.
// ~1e-16 of max|U| (the parity tolerance is 1e-13).
// (E holds the x table followed by the y table: one launch for both)
__global__ void ackley_table_kernel(const double* __restrict__ ux, int n_ux, const double* __restrict__ uy, int n_uy,
                                    const double* __restrict__ mu, int64_t n_s, int n_mu, double* __restrict__ E) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)(n_ux + n_uy) * n_s) return;
    const int iu = (int)(idx / n_s); const int64_t s = idx % n_s;
    const double w = mul(6.283185307179586, add(1., mul(.1, mu[s * n_mu + 0])));
    E[idx] = exp(mul(.5, cos(mul(w, iu < n_ux ? ux[iu] : uy[iu - n_ux]))));
}

__global__ void ackley_indexed_kernel(const double* __restrict__ X, int64_t n_xyz, int64_t n_rows, int64_t row0,
                                      const int* __restrict__ ix, const int* __restrict__ iy,
                                      const double* __restrict__ Ex, const double* __restrict__ Ey,
                                      const double* __restrict__ mu, int64_t n_s, int n_mu,
                                      double* __restrict__ U, int64_t ldu) {
    const int64_t col = (int64_t)blockIdx.y * COLS_PER_BLOCK + threadIdx.x;
    const int64_t r_begin = (int64_t)blockIdx.x * NODES_PER_BLOCK;
    const int n_here = (int)(min(n_rows, r_begin + NODES_PER_BLOCK) - r_begin);
    __shared__ double sr[NODES_PER_BLOCK];
    __shared__ int64_t sxo[NODES_PER_BLOCK];          // row offset of the node's x entry in the table (ix * n_s)
    __shared__ int sy[NODES_PER_BLOCK];
    for (int e = threadIdx.x; e < n_here; e += blockDim.x) {
        const int64_t r = r_begin + e;
        double x = X[row0 + r], y = X[n_xyz + row0 + r];
        sr[e] = sqrt(mul(.5, add(mul(x, x), mul(y, y))));
        sxo[e] = (int64_t)ix[row0 + r] * n_s; sy[e] = iy[row0 + r];
    }
    __syncthreads();
    if (col >= n_s) return;
    const double m1 = mu[col * n_mu + 1], m2 = mu[col * n_mu + 2];
    const double a2 = mul(-20., add(1., mul(.1, m2)));
    const double c1 = mul(-.2, add(1., mul(.1, m1)));
    const double e1 = 2.718281828459045;
    // the loop is bound by instruction issue (exp is ~35 instructions): keep the index arithmetic to one pointer
    // bump and one table offset per element
    const double* __restrict__ exc = Ex + col;
    const double* __restrict__ eyc = Ey + col;
    double* __restrict__ up = U + r_begin * ldu + col;
    int last_y = -1; double ey = 0.;          // consecutive nodes of a tensor grid share y: one Ey load per grid row
#pragma unroll 4
    for (int e = 0; e < n_here; ++e, up += ldu) {
        const double t1 = mul(a2, exp(mul(c1, sr[e])));
        if (sy[e] != last_y) { last_y = sy[e]; ey = eyc[(int64_t)last_y * n_s]; }
        const double t2 = mul(exc[sxo[e]], ey);
        *up = add(add(sub(t1, t2), 20.), e1);
    }
}

// Ackley on radius-sorted nodes.  The first term a2 exp(c1 r) only depends on (r, sample), and a symmetric mesh
// repeats radii (the 400 x 400 grid of C2 has 31 965 distinct values of r for 160 000 nodes): the host sorts the
// nodes of this call by r (perm = node ids, rs = their radii), a thread keeps the last exp and re-evaluates it
// only when r changes -- a warp-uniform branch -- so ~1 exp per 5 elements is left and the kernel is bound by
// the HBM writes.  A row is still written as one contiguous segment per CTA; bit-identical to ackley_indexed_kernel.
template <int VEC>   // VEC = 2: two adjacent columns per thread (16-byte table loads and stores; n_s, ldu even)
__global__ void ackley_grouped_kernel(const int* __restrict__ perm, const double* __restrict__ rs, int64_t n_sorted,
                                      int64_t row0, const int* __restrict__ ix, const int* __restrict__ iy,
                                      const double* __restrict__ Ex, const double* __restrict__ Ey,
                                      const double* __restrict__ mu, int64_t n_s, int n_mu,
                                      double* __restrict__ U, int64_t ldu) {
    const int64_t col = ((int64_t)blockIdx.y * COLS_PER_BLOCK + threadIdx.x) * VEC;
    const int64_t e_begin = (int64_t)blockIdx.x * NODES_PER_BLOCK;
    const int n_here = (int)(min(n_sorted, e_begin + NODES_PER_BLOCK) - e_begin);
    __shared__ double sr[NODES_PER_BLOCK];
    __shared__ int4 sidx[NODES_PER_BLOCK];           // {ix, iy, local row, -}: one LDS.128 per node
    for (int e = threadIdx.x; e < n_here; e += blockDim.x) {
        const int node = perm[e_begin + e];
        sr[e] = rs[e_begin + e];
        sidx[e] = make_int4(ix[node], iy[node], (int)(node - row0), 0);
    }
    __syncthreads();
    if (col >= n_s) return;
    double a2[VEC], c1[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int64_t c = min(col + v, n_s - 1);
        a2[v] = mul(-20., add(1., mul(.1, mu[c * n_mu + 2])));
        c1[v] = mul(-.2, add(1., mul(.1, mu[c * n_mu + 1])));
    }
    const double e1 = 2.718281828459045;
    const double* __restrict__ exc = Ex + col;
    const double* __restrict__ eyc = Ey + col;
    double* __restrict__ uc = U + col;
    const int ns = (int)n_s, ld = (int)ldu;
    double last_r = -1., t1[VEC] = {};
#pragma unroll 4
    for (int e = 0; e < n_here; ++e) {
        const int4 id = sidx[e];
        double ex[VEC], ey[VEC], out[VEC];
        if (VEC == 2) {
            const double2 a = *reinterpret_cast<const double2*>(exc + (int64_t)id.x * ns);
            const double2 b = *reinterpret_cast<const double2*>(eyc + (int64_t)id.y * ns);
            ex[0] = a.x; ex[VEC - 1] = a.y; ey[0] = b.x; ey[VEC - 1] = b.y;
        } else {
            ex[0] = exc[(int64_t)id.x * ns]; ey[0] = eyc[(int64_t)id.y * ns];
        }
        const double r = sr[e];
        if (r != last_r) {
            last_r = r;
#pragma unroll
            for (int v = 0; v < VEC; ++v) t1[v] = mul(a2[v], exp(mul(c1[v], r)));
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = add(add(sub(t1[v], mul(ex[v], ey[v])), 20.), e1);
        double* dst = uc + (int64_t)id.z * ld;
        // streaming stores: U is written once and must not push the two tables out of L2
        if (VEC == 2) __stcs(reinterpret_cast<double2*>(dst), make_double2(out[0], out[VEC - 1]));
        else __stcs(dst, out[0]);
    }
}

// Burgers, experiments/1dt_burger/hyperparams.py:45-55.  Column = i_sample * n_t + j_time.
__global__ void burgers_kernel(const double* __restrict__ X, int64_t n_rows, int64_t row0,
                               const double* __restrict__ mu, int64_t n_s, int n_mu,
                               const double* __restrict__ t, int n_t,
                               double* __restrict__ U, int64_t ldu, int npb) {
    const int64_t n_st = n_s * n_t;
    const int64_t col = (int64_t)blockIdx.y * COLS_PER_BLOCK + threadIdx.x;
    const int64_t r_begin = (int64_t)blockIdx.x * npb;
    const int64_t r_end = min(n_rows, r_begin + npb);
    if (col >= n_st) return;
    const int64_t i = col / n_t; const int j = (int)(col % n_t);
    const double m = mu[i * n_mu], tj = t[j];
    const double four_mu = mul(4., m);
    if (tj == 1.) {
        const double k = dvd(1., four_mu);                         // 1/(4*mu)
        for (int64_t r = r_begin; r < r_end; ++r) {
            double x = X[row0 + r];
            double ex = exp(mul(k, sub(mul(x, x), .25)));          // exp(1/(4*mu)*(x**2 - 1/4))
            U[r * ldu + col] = dvd(x, add(1., ex));
        }
    } else {
        const double t0 = exp(dvd(1., mul(8., m)));                // exp(1/(8*mu))
        const double sq = sqrt(dvd(tj, t0));                       // sqrt(t/t0)
        const double den = mul(four_mu, tj);                       // 4*mu*t
        for (int64_t r = r_begin; r < r_end; ++r) {
            double x = X[row0 + r];
            double ex = exp(dvd(mul(x, x), den));                  // exp(x**2/(4*mu*t)), inf allowed
            U[r * ldu + col] = dvd(dvd(x, tj), add(1., mul(sq, ex)));
        }
    }
}

// DCT-II factor tables for the known-answer matrix: out[i][k] = scale_k * sqrt(2/n) cos(pi (i0+i+1/2)(k+1)/n)
__global__ void dct_table_kernel(double* __restrict__ out, int64_t rows, int r, int64_t i0, int64_t n,
                                 double decay, int transposed, int64_t ld) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * r) return;
    int64_t i = idx / r; int k = (int)(idx % r);
    // exact argument reduction: (2i+1)(k+1) mod 4n in integers, then cospi of a value in [0, 2)
    int64_t num = ((2 * (i0 + i) + 1) % (4 * n)) * (int64_t)(k + 1) % (4 * n);
    double c = cospi((double)num / (double)(2 * n));
    double v = sqrt(2. / (double)n) * c;
    if (decay != 0.) v *= exp2(-decay * (double)k);
    if (transposed) out[(int64_t)k * ld + i] = v; else out[i * ld + k] = v;
}

}  // namespace

extern "C" int pb_snapshots(pb_ctx* ctx, int field, int dim, int64_t n_xyz, const double* X_dev,
                            int64_t n_s, int n_mu, const double* mu_dev, int n_t, const double* t_dev,
                            int64_t row0, int64_t n_rows, double* U_dev, int64_t ldu) {
    PB_ENTER(ctx);
    if (n_xyz <= 0 || n_s <= 0 || n_rows < 0 || row0 < 0 || row0 + n_rows > n_xyz)
        PB_FAIL(ctx, PB_ERR_SHAPE, "pb_snapshots: bad node range [%lld, %lld) of %lld", (long long)row0,
                (long long)(row0 + n_rows), (long long)n_xyz);
    const int64_t n_st = n_s * (n_t > 0 ? n_t : 1);
    if (ldu < n_st) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_snapshots: ldu %lld < n_st %lld", (long long)ldu, (long long)n_st);
    if (n_rows == 0) return PB_OK;
    cudaEventRecord(ctx->ev0, ctx->stream);
    dim3 grid((unsigned)((n_rows + NODES_PER_BLOCK - 1) / NODES_PER_BLOCK),
              (unsigned)((n_st + COLS_PER_BLOCK - 1) / COLS_PER_BLOCK));
    if (grid.y > 65535) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_snapshots: too many snapshot columns (%lld)", (long long)n_st);
    // Shekel / Burgers are bound by fp64 issue (divisions, exp) on small meshes: fewer nodes per CTA until the grid
    // holds >= 8 CTAs of 128 threads per SM (C1: 400 nodes x 1000 samples was 56 CTAs on 148 SMs)
    int npb = NODES_PER_BLOCK;
    while (npb > 2 && ((n_rows + npb - 1) / npb) * (int64_t)grid.y < 8 * (int64_t)ctx->sm_count) npb /= 2;
    dim3 grid_small((unsigned)((n_rows + npb - 1) / npb), grid.y);
    switch (field) {
    case PB_FIELD_SHEKEL:
        if (dim != 1 || n_t > 0 || n_mu < 2 || (n_mu % 2)) PB_FAIL(ctx, PB_ERR_SHAPE, "shekel: dim=1, steady, even n_mu");
        shekel_kernel<<<grid_small, COLS_PER_BLOCK, (size_t)n_mu * COLS_PER_BLOCK * sizeof(double), ctx->stream>>>(
            X_dev, n_rows, row0, mu_dev, n_s, n_mu, U_dev, ldu, npb);
        break;
    case PB_FIELD_ACKLEY:
        if (dim != 2 || n_t > 0 || n_mu != 3) PB_FAIL(ctx, PB_ERR_SHAPE, "ackley: dim=2, steady, n_mu=3");
        ackley_kernel<<<grid, COLS_PER_BLOCK, 0, ctx->stream>>>(X_dev, n_xyz, n_rows, row0, mu_dev, n_s, n_mu, U_dev, ldu);
        break;
    case PB_FIELD_BURGERS:
        if (dim != 1 || n_t <= 0 || n_mu != 1 || !t_dev) PB_FAIL(ctx, PB_ERR_SHAPE, "burgers: dim=1, n_t>0, n_mu=1");
        burgers_kernel<<<grid_small, COLS_PER_BLOCK, 0, ctx->stream>>>(X_dev, n_rows, row0, mu_dev, n_s, n_mu, t_dev, n_t, U_dev, ldu, npb);
        break;
    default:
        PB_FAIL(ctx, PB_ERR_UNSUPPORTED_FIELD, "pb_snapshots: field id %d is not registered (no CPU fallback)", field);
    }
    PB_LAUNCH_CHECK(ctx);
    cudaEventRecord(ctx->ev1, ctx->stream);
    PB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->ms[PB_T_SNAPSHOTS] = ms;
    return PB_OK;
}

extern "C" int pb_snapshots_ackley_indexed(pb_ctx* ctx, int64_t n_xyz, const double* X_dev, int n_ux, const double* ux_dev,
                                           const int* ix_dev, int n_uy, const double* uy_dev, const int* iy_dev,
                                           int64_t n_s, const double* mu_dev, int64_t row0, int64_t n_rows,
                                           double* U_dev, int64_t ldu) {
    PB_ENTER(ctx);
    if (n_xyz <= 0 || n_s <= 0 || n_rows < 0 || row0 < 0 || row0 + n_rows > n_xyz || n_ux <= 0 || n_uy <= 0 || ldu < n_s)
        PB_FAIL(ctx, PB_ERR_SHAPE, "pb_snapshots_ackley_indexed: bad shape");
    if (n_rows == 0) return PB_OK;
    cudaEventRecord(ctx->ev0, ctx->stream);
    PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)(n_ux + n_uy) * n_s));
    double* Ex = ctx->ws; double* Ey = Ex + (size_t)n_ux * n_s;
    ackley_table_kernel<<<(unsigned)(((int64_t)(n_ux + n_uy) * n_s + 255) / 256), 256, 0, ctx->stream>>>(ux_dev, n_ux, uy_dev, n_uy,
                                                                                                        mu_dev, n_s, 3, Ex);
    PB_LAUNCH_CHECK(ctx);
    dim3 grid((unsigned)((n_rows + NODES_PER_BLOCK - 1) / NODES_PER_BLOCK), (unsigned)((n_s + COLS_PER_BLOCK - 1) / COLS_PER_BLOCK));
    if (grid.y > 65535) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_snapshots_ackley_indexed: too many snapshot columns");
    ackley_indexed_kernel<<<grid, COLS_PER_BLOCK, 0, ctx->stream>>>(X_dev, n_xyz, n_rows, row0, ix_dev, iy_dev, Ex, Ey, mu_dev, n_s, 3,
                                                                  U_dev, ldu);
    PB_LAUNCH_CHECK(ctx);
    cudaEventRecord(ctx->ev1, ctx->stream);
    PB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->ms[PB_T_SNAPSHOTS] = ms;
    return PB_OK;
}

extern "C" int pb_snapshots_ackley_grouped(pb_ctx* ctx, int64_t n_xyz, int n_ux, const double* ux_dev, const int* ix_dev,
                                           int n_uy, const double* uy_dev, const int* iy_dev, int64_t n_s,
                                           const double* mu_dev, int64_t row0, int64_t n_rows, const int* perm_dev,
                                           const double* r_sorted_dev, double* U_dev, int64_t ldu) {
    PB_ENTER(ctx);
    if (n_xyz <= 0 || n_s <= 0 || n_rows < 0 || row0 < 0 || row0 + n_rows > n_xyz || n_ux <= 0 || n_uy <= 0 || ldu < n_s ||
        n_xyz > INT32_MAX)
        PB_FAIL(ctx, PB_ERR_SHAPE, "pb_snapshots_ackley_grouped: bad shape");
    if (n_rows == 0) return PB_OK;
    cudaEventRecord(ctx->ev0, ctx->stream);
    PB_TRY(pb_reserve(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)(n_ux + n_uy) * n_s));
    double* Ex = ctx->ws; double* Ey = Ex + (size_t)n_ux * n_s;
    ackley_table_kernel<<<(unsigned)(((int64_t)(n_ux + n_uy) * n_s + 255) / 256), 256, 0, ctx->stream>>>(ux_dev, n_ux, uy_dev, n_uy,
                                                                                                        mu_dev, n_s, 3, Ex);
    PB_LAUNCH_CHECK(ctx);
    const bool vec2 = n_s % 2 == 0 && ldu % 2 == 0 && ((uintptr_t)U_dev % 16) == 0 && ldu < INT32_MAX;
    const int cols_per_block = COLS_PER_BLOCK * (vec2 ? 2 : 1);
    dim3 grid((unsigned)((n_rows + NODES_PER_BLOCK - 1) / NODES_PER_BLOCK), (unsigned)((n_s + cols_per_block - 1) / cols_per_block));
    if (grid.y > 65535 || (int64_t)(n_ux > n_uy ? n_ux : n_uy) * n_s > INT32_MAX || ldu > INT32_MAX)
        PB_FAIL(ctx, PB_ERR_SHAPE, "pb_snapshots_ackley_grouped: too many snapshot columns");
    if (vec2)
        ackley_grouped_kernel<2><<<grid, COLS_PER_BLOCK, 0, ctx->stream>>>(perm_dev, r_sorted_dev, n_rows, row0, ix_dev, iy_dev, Ex, Ey,
                                                                         mu_dev, n_s, 3, U_dev, ldu);
    else
        ackley_grouped_kernel<1><<<grid, COLS_PER_BLOCK, 0, ctx->stream>>>(perm_dev, r_sorted_dev, n_rows, row0, ix_dev, iy_dev, Ex, Ey,
                                                                         mu_dev, n_s, 3, U_dev, ldu);
    PB_LAUNCH_CHECK(ctx);
    cudaEventRecord(ctx->ev1, ctx->stream);
    PB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->ms[PB_T_SNAPSHOTS] = ms;
    return PB_OK;
}

// Closed-form left basis of the known-answer matrix: Phi[i][k] = sqrt(2/n_h) cos(pi (row0 + i + 1/2)(k + 1) / n_h)
extern "C" int pb_dct_basis(pb_ctx* ctx, int64_t n_h, int r, int64_t row0, int64_t n_rows, double* Phi_dev, int64_t ld) {
    PB_ENTER(ctx);
    if (r <= 0 || ld < r || n_rows <= 0 || row0 < 0 || row0 + n_rows > n_h) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_dct_basis: bad shape");
    const int64_t ne = n_rows * r;
    dct_table_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(Phi_dev, n_rows, r, row0, n_h, 0., 0, ld);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

extern "C" int pb_dct_lowrank(pb_ctx* ctx, int64_t n_h, int64_t n_s, int r, double decay,
                              int64_t row0, int64_t n_rows, double* U_dev, int64_t ldu) {
    PB_ENTER(ctx);
    if (r <= 0 || r > n_s || row0 + n_rows > n_h) PB_FAIL(ctx, PB_ERR_SHAPE, "pb_dct_lowrank: bad shape");
    double *phi = nullptr, *psit = nullptr;
    const int64_t chunk = 1 << 20;                      // rows per GEMM call (bounds the phi table)
    PB_CUDA(ctx, cudaMalloc(&phi, sizeof(double) * (size_t)std::min<int64_t>(chunk, n_rows) * r));
    PB_CUDA(ctx, cudaMalloc(&psit, sizeof(double) * (size_t)r * n_s));
    int64_t n = (int64_t)r * n_s;
    dct_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(psit, n_s, r, 0, n_s, 0., 1, n_s);
    PB_LAUNCH_CHECK(ctx);
    int rc = PB_OK;
    for (int64_t h = 0; h < n_rows && rc == PB_OK; h += chunk) {
        int64_t rows = std::min(chunk, n_rows - h);
        int64_t ne = rows * r;
        dct_table_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(phi, rows, r, row0 + h, n_h, decay, 0, r);
        ctx->launches++;
        rc = pbi_gemm_nn(ctx, phi, r, rows, r, psit, n_s, n_s, U_dev + h * ldu, ldu, nullptr, 0, nullptr);
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(phi); cudaFree(psit);
    return rc;
}
