// Device helpers shared by the DMMA kernels: the fp64 tensor instruction, cp.async staging
// and the zero-filling tile loader.  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pbt {

constexpr int BK = 16;          // K-chunk (rows of the reduction index) per pipeline stage
constexpr int NTHREADS = 256;   // every tiled kernel runs 8 warps per CTA

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(s), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" :: "n"(N));
}

// Copy a (R x C) tile of a row-major matrix (row pitch ld) whose top-left element is
// (r0, c0) into shared memory with pitch P; rows >= r_end and columns >= c_end are zero-filled.
template <int R, int C, int P, bool ALIGN16>
__device__ __forceinline__ void load_tile(double* __restrict__ s, const double* __restrict__ g,
                                          int64_t ld, int64_t r0, int64_t r_end,
                                          int64_t c0, int64_t c_end, int tid) {
    if (ALIGN16) {
        constexpr int CH = C / 2;                 // 16-byte chunks per row
        constexpr int TOTAL = R * CH;
#pragma unroll
        for (int q = 0; q < (TOTAL + NTHREADS - 1) / NTHREADS; ++q) {
            int c = tid + q * NTHREADS;
            if (TOTAL % NTHREADS != 0 && c >= TOTAL) break;
            int row = c / CH, cc = (c % CH) * 2;
            int64_t gr = r0 + row, gc = c0 + cc;
            int bytes = 0;
            if (gr < r_end) { if (gc + 1 < c_end) bytes = 16; else if (gc < c_end) bytes = 8; }
            const double* src = bytes ? g + gr * ld + gc : g;
            cp_async16(s + row * P + cc, src, bytes);
        }
    } else {
        constexpr int TOTAL = R * C;
#pragma unroll
        for (int q = 0; q < (TOTAL + NTHREADS - 1) / NTHREADS; ++q) {
            int c = tid + q * NTHREADS;
            if (TOTAL % NTHREADS != 0 && c >= TOTAL) break;
            int row = c / C, cc = c % C;
            int64_t gr = r0 + row, gc = c0 + cc;
            int bytes = (gr < r_end && gc < c_end) ? 8 : 0;
            const double* src = bytes ? g + gr * ld + gc : g;
            cp_async8(s + row * P + cc, src, bytes);
        }
    }
}

}  // namespace pbt
