// awegpu_reduce.cu — device-resident products of a batch of constraint Jacobians (first slice).
//
// awegpu_eval_batch leaves G [B, n_g] and the CCS values Jvals [B, nnz] of every instance in HBM.  A batched solver
// (the natural sequel to awebox/sweep.py:117-204 and awebox/pmpc.py:221-302) does not need the O(nnz) values on the
// host: it needs products and norms, O(n_v + n_g) per instance.  One kernel, one pass over Jvals:
//     JTl   [B, n_v] = J_b^T lambda_b          (dual feasibility residual / Lagrangian gradient part)
//     Jv    [B, n_g] = J_b v_b                 (linearised constraint change along a step)
//     cnorm [B, n_v] = max_i |J_b[i, c]|,  rnorm [B, n_g] = max_c |J_b[r, c]|     (scaling)
//     ginf  [B]      = max_i |g_b[i]|
// Thread = one column of one instance: it walks the contiguous entries of its column, so J^T lambda and the column
// norm stay in registers; J v and the row norms are accumulated with reductions in L2 (red.global.add.f64 /
// red.global.max.u64, no return value: the thread does not wait for them; non-negative doubles order like their bit
// patterns).  A first version accumulated them in shared memory, one CTA per instance: FP64 shared-memory atomics are
// compare-and-swap loops, and that kernel took 23 ms for 4096 AP2 instances.  HBM-bound: Jvals is read exactly once (8 nnz
// bytes per instance), everything else is L1/L2 traffic.
#include <cuda_runtime.h>
#include <cstdint>

#define RED_THREADS 256

#include "awegpu_reduce.h"

__global__ void __launch_bounds__(RED_THREADS) awegpu_reduce_kernel(const ReduceArgs a)
{
    __shared__ double s_red[RED_THREADS / 32];
    const int tid = threadIdx.x;
    const bool want_y = a.Jv != nullptr && a.v != nullptr, want_rn = a.rnorm != nullptr;
    const int col_blocks = (a.n_v + RED_THREADS - 1) / RED_THREADS;
    const long long n_work = (long long)a.B * col_blocks;
    for (long long w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int b = (int)(w / col_blocks), cb = (int)(w - (long long)b * col_blocks);
        const int c = cb * RED_THREADS + tid;
        const double *Jb = a.J + (size_t)b * (size_t)a.nnz;
        const double *lam = a.lambda ? a.lambda + (size_t)b * a.ld_lambda : nullptr;
        if (c < a.n_v) {
            const int e0 = __ldg(a.colind + c), e1 = __ldg(a.colind + c + 1);
            const double vc = want_y ? __ldg(a.v + (size_t)b * a.ld_v + c) : 0.0;
            double *yb = want_y ? a.Jv + (size_t)b * a.n_g : nullptr;
            unsigned long long *rb = want_rn ? reinterpret_cast<unsigned long long *>(a.rnorm + (size_t)b * a.n_g) : nullptr;
            double acc = 0.0, cn = 0.0;
            for (int e = e0; e < e1; ++e) {
                const double x = Jb[e];
                const int r = __ldg(a.row + e);
                if (lam) acc = fma(x, __ldg(lam + r), acc);
                if (want_y) atomicAdd(yb + r, x * vc);
                const double ax = fabs(x);
                cn = fmax(cn, ax);
                if (want_rn) atomicMax(rb + r, (unsigned long long)__double_as_longlong(ax));
            }
            if (a.JTl && lam) a.JTl[(size_t)b * a.n_v + c] = acc;
            if (a.cnorm) a.cnorm[(size_t)b * a.n_v + c] = cn;
        }
        if (a.ginf && cb == 0) {
            double m = 0.0;
            for (int i = tid; i < a.n_g; i += RED_THREADS) m = fmax(m, fabs(a.G[(size_t)b * a.n_g + i]));
            for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((tid & 31) == 0) s_red[tid >> 5] = m;
            __syncthreads();
            if (tid == 0) {
                m = 0.0;
                for (int q = 0; q < RED_THREADS / 32; ++q) m = fmax(m, s_red[q]);
                a.ginf[b] = m;
            }
            __syncthreads();
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
    if (e != cudaSuccess) return (int)e;
    const long long n_work = (long long)a.B * ((a.n_v + RED_THREADS - 1) / RED_THREADS);
    long long grid = (long long)n_sm * 8;
    if (grid > n_work) grid = n_work;
    if (grid < 1) return 0;
    awegpu_reduce_kernel<<<(int)grid, RED_THREADS, 0, st>>>(a);
    return (int)cudaGetLastError();
}
