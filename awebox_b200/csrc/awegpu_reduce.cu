// awegpu_reduce.cu — device-resident products of a batch of constraint Jacobians (first slice).
//
// awegpu_eval_batch leaves G [B, n_g] and the CCS values Jvals [B, nnz] of every instance in HBM.  A batched solver
// (the natural sequel to awebox/sweep.py:117-204 and awebox/pmpc.py:221-302) does not need the O(nnz) values on the
// host: it needs products and norms, O(n_v + n_g) per instance.  One kernel, one pass over Jvals:
//     JTl   [B, n_v] = J_b^T lambda_b          (dual feasibility residual / Lagrangian gradient part)
//     Jv    [B, n_g] = J_b v_b                 (linearised constraint change along a step)
//     cnorm [B, n_v] = max_i |J_b[i, c]|,  rnorm [B, n_g] = max_c |J_b[r, c]|     (scaling)
//     ginf  [B]      = max_i |g_b[i]|
// Warp = 32 consecutive CCS entries of one instance (coalesced reads of the values, rows and columns): J v and the row
// norms go to L2 as reductions, one per entry (red.global.add.f64 / red.global.max.u64 on the bit pattern of the
// magnitude, no return value: nobody waits for them); J^T lambda and the column norms are first reduced inside the warp
// by a segmented scan over the column index (entries of a column are consecutive), so that only the last lane of every
// column segment issues a reduction.  HBM-bound by design: the values are read exactly once (8 nnz bytes per instance).
// History: thread = column with shared-memory accumulators, one CTA per instance: 23 ms for 4096 AP2 instances (FP64
// shared atomics are compare-and-swap loops); thread = column with L2 reductions: 6.1 ms (every lane walks its own
// column: 32 sectors per load instruction); this version: see profiles/.
#include <cuda_runtime.h>
#include <cstdint>

#define RED_THREADS 256

#include "awegpu_reduce.h"

__global__ void __launch_bounds__(RED_THREADS) awegpu_reduce_kernel(const ReduceArgs a)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * RED_THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * RED_THREADS) >> 5;
    const int chunks = (int)((a.nnz + 31) / 32);                 // per instance
    const bool want_y = a.Jv != nullptr && a.v != nullptr, want_rn = a.rnorm != nullptr;
    const bool want_jtl = a.JTl != nullptr && a.lambda != nullptr, want_cn = a.cnorm != nullptr;
    for (long long w = warp; w < (long long)a.B * chunks; w += n_warps) {
        const int b = (int)(w / chunks);
        const long long e = (long long)(w - (long long)b * chunks) * 32 + lane;
        const bool on = e < a.nnz;
        double x = 0.0;
        int r = 0, c = -1;
        if (on) {
            x = a.J[(size_t)b * (size_t)a.nnz + e];
            r = __ldg(a.row + e), c = __ldg(a.col + e);
        }
        const double ax = fabs(x);
        if (on && want_y) atomicAdd(a.Jv + (size_t)b * a.n_g + r, x * __ldg(a.v + (size_t)b * a.ld_v + c));
        if (on && want_rn)
            atomicMax(reinterpret_cast<unsigned long long *>(a.rnorm + (size_t)b * a.n_g) + r, (unsigned long long)__double_as_longlong(ax));
        if (want_jtl || want_cn) {
            double t = (on && want_jtl) ? x * __ldg(a.lambda + (size_t)b * a.ld_lambda + r) : 0.0;
            double m = ax;
            // segmented inclusive scan over the column index
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double tu = __shfl_up_sync(0xffffffffu, t, d), mu = __shfl_up_sync(0xffffffffu, m, d);
                const int cu = __shfl_up_sync(0xffffffffu, c, d);
                if (lane >= d && cu == c) t += tu, m = fmax(m, mu);
            }
            const int cn = __shfl_down_sync(0xffffffffu, c, 1);
            if (on && (lane == 31 || cn != c)) {
                if (want_jtl) atomicAdd(a.JTl + (size_t)b * a.n_v + c, t);
                if (want_cn)
                    atomicMax(reinterpret_cast<unsigned long long *>(a.cnorm + (size_t)b * a.n_v) + c, (unsigned long long)__double_as_longlong(m));
            }
        }
    }
    if (a.ginf) {
        // max |g| per instance: one warp per instance
        for (long long b = warp; b < a.B; b += n_warps) {
            double m = 0.0;
            for (int i = lane; i < a.n_g; i += 32) m = fmax(m, fabs(a.G[(size_t)b * a.n_g + i]));
            for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) a.ginf[b] = m;
        }
    }
}

// 0 = launched; > 0 = cudaError_t
int awe_reduce_launch(cudaStream_t st, const ReduceArgs &a, int n_sm)
{
    cudaError_t e = cudaSuccess;
    // accumulated with reductions: start from zero (0.0 is also the identity of the max of magnitudes)
    if (a.Jv && a.v) e = cudaMemsetAsync(a.Jv, 0, (size_t)a.B * a.n_g * sizeof(double), st);
    if (e == cudaSuccess && a.rnorm) e = cudaMemsetAsync(a.rnorm, 0, (size_t)a.B * a.n_g * sizeof(double), st);
    if (e == cudaSuccess && a.JTl && a.lambda) e = cudaMemsetAsync(a.JTl, 0, (size_t)a.B * a.n_v * sizeof(double), st);
    if (e == cudaSuccess && a.cnorm) e = cudaMemsetAsync(a.cnorm, 0, (size_t)a.B * a.n_v * sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
    const long long n_work = (long long)a.B * ((a.nnz + 31) / 32);
    long long grid = (long long)n_sm * 8;
    if (grid * (RED_THREADS / 32) > n_work) grid = (n_work + RED_THREADS / 32 - 1) / (RED_THREADS / 32);
    if (grid < 1) return 0;
    awegpu_reduce_kernel<<<(int)grid, RED_THREADS, 0, st>>>(a);
    return (int)cudaGetLastError();
}
