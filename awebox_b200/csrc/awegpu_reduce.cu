// awegpu_reduce.cu — device-resident products of a batch of constraint Jacobians (first slice).
//
// awegpu_eval_batch leaves G [B, n_g] and the CCS values Jvals [B, nnz] of every instance in HBM.  A batched solver
// (the natural sequel to awebox/sweep.py:117-204 and awebox/pmpc.py:221-302) does not need the O(nnz) values on the
// host: it needs products and norms, O(n_v + n_g) per instance.  One kernel, one pass over Jvals:
//     JTl   [B, n_v] = J_b^T lambda_b          (dual feasibility residual / Lagrangian gradient part)
//     Jv    [B, n_g] = J_b v_b                 (linearised constraint change along a step)
//     cnorm [B, n_v] = max_i |J_b[i, c]|,  rnorm [B, n_g] = max_c |J_b[r, c]|     (scaling)
//     ginf  [B]      = max_i |g_b[i]|
// CTA = one instance at a time (persistent over the batch), thread = column: a thread walks the contiguous entries of
// its column, so J^T lambda and the column norm stay in a register; J v and the row norm are scattered into
// shared-memory accumulators (n_g doubles + n_g 64-bit keys).  HBM-bound: Jvals is read exactly once (8 nnz bytes per
// instance), everything else is L1/L2 traffic.
#include <cuda_runtime.h>
#include <cstdint>

#define RED_THREADS 256

#include "awegpu_reduce.h"

__global__ void __launch_bounds__(RED_THREADS) awegpu_reduce_kernel(const ReduceArgs a)
{
    extern __shared__ double red_smem[];
    double *y = red_smem;                                                    // [n_g]  J v
    unsigned long long *rn = reinterpret_cast<unsigned long long *>(y + a.n_g);   // [n_g]  bit pattern of max |J| per row
    __shared__ double s_red[RED_THREADS / 32];
    const int tid = threadIdx.x;
    const bool want_y = a.Jv != nullptr && a.v != nullptr, want_rn = a.rnorm != nullptr;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        if (want_y || want_rn) {
            for (int i = tid; i < a.n_g; i += RED_THREADS) y[i] = 0.0, rn[i] = 0ull;
            __syncthreads();
        }
        const double *Jb = a.J + (size_t)b * (size_t)a.nnz;
        const double *lam = a.lambda ? a.lambda + (size_t)b * a.ld_lambda : nullptr;
        const double *vv = a.v ? a.v + (size_t)b * a.ld_v : nullptr;
        for (int c = tid; c < a.n_v; c += RED_THREADS) {
            const int e0 = __ldg(a.colind + c), e1 = __ldg(a.colind + c + 1);
            const double vc = vv ? __ldg(vv + c) : 0.0;
            double acc = 0.0, cn = 0.0;
            for (int e = e0; e < e1; ++e) {
                const double x = Jb[e];
                const int r = __ldg(a.row + e);
                if (lam) acc = fma(x, __ldg(lam + r), acc);
                if (want_y) atomicAdd(y + r, x * vc);
                const double ax = fabs(x);
                cn = fmax(cn, ax);
                // non-negative doubles order like their bit patterns
                if (want_rn) atomicMax(rn + r, (unsigned long long)__double_as_longlong(ax));
            }
            if (a.JTl && lam) a.JTl[(size_t)b * a.n_v + c] = acc;
            if (a.cnorm) a.cnorm[(size_t)b * a.n_v + c] = cn;
        }
        if (a.ginf) {
            double m = 0.0;
            for (int i = tid; i < a.n_g; i += RED_THREADS) m = fmax(m, fabs(a.G[(size_t)b * a.n_g + i]));
            for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((tid & 31) == 0) s_red[tid >> 5] = m;
        }
        __syncthreads();
        if (a.ginf && tid == 0) {
            double m = 0.0;
            for (int w = 0; w < RED_THREADS / 32; ++w) m = fmax(m, s_red[w]);
            a.ginf[b] = m;
        }
        if (want_y)
            for (int i = tid; i < a.n_g; i += RED_THREADS) a.Jv[(size_t)b * a.n_g + i] = y[i];
        if (want_rn)
            for (int i = tid; i < a.n_g; i += RED_THREADS) a.rnorm[(size_t)b * a.n_g + i] = __longlong_as_double((long long)rn[i]);
        __syncthreads();
    }
}

// -> cudaError_t as int; smem_needed is returned through *smem_bytes when it exceeds what an SM offers (error 1)
int awe_reduce_launch(cudaStream_t st, const ReduceArgs &a, int n_sm)
{
    const size_t smem = (size_t)a.n_g * 16;
    if (smem > 220 * 1024) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(awegpu_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int per_sm = smem > 0 ? (int)((220 * 1024) / smem) : 8;
    int grid = n_sm * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    if (grid > a.B) grid = a.B;
    if (grid < 1) return 0;
    awegpu_reduce_kernel<<<grid, RED_THREADS, smem, st>>>(a);
    return (int)cudaGetLastError();
}
