// awegpu_out.cu — outputs map of the OCP: model.outputs_fun over the shooting and collocation
// nodes of every interval, what the reference builds as Outputs_fun(V, P) (awebox/ocp/collocation.py:366-381,
// awebox/ocp/discretization.py:316-334) and reads by label after every solve.  Expressions: awebox_b200/mdl/outputs.py.
//
//   awegpu_out_kernel   thread = one node of one instance, consecutive lanes = consecutive nodes of the instance
//                       (node order of the reference: for k: shooting node k, collocation nodes (k, 0..d-1)).
//                       Node inputs are gathered straight from the instance's row of W (the interval slab is contiguous
//                       in V; a warp's nodes cover ~6 neighbouring slabs, served by L1): collocation nodes take xdot
//                       from the collocation polynomial, shooting nodes the lifted V[xdot, k].  The generated
//                       straight-line program (node_out.inc) stores every entry as it becomes final into
//                       Out[b][entry][node]: lanes write consecutive addresses (whole 256-byte runs, no staging, no
//                       read-modify-write), and one label of one instance is a contiguous time series — the access
//                       pattern of the reference's consumers.  Write-bound: 8 n_out bytes per node against
//                       8 (n_v / n_nodes) bytes read.
#include <cuda_runtime.h>
#include <cstdint>

#include "awegpu_out.h"
#include "model_defs.h"
#include "out_defs.h"
#include "awegpu_arith.cuh"

#define N_X AWE_N_X
#define N_U AWE_N_U
#define N_Z AWE_N_Z
#define N_TH AWE_N_THETA
#define N_D AWE_D
#define OFF_U (N_X)
#define OFF_XDOT (N_X + N_U)
#define OFF_Z (2 * N_X + N_U)
#define OFF_CX(j) (2 * N_X + N_U + N_Z + (j) * (N_X + N_Z))
#define OFF_CZ(j) (OFF_CX(j) + N_X)
#define SLAB_W (2 * N_X + N_U + N_Z + N_D * (N_X + N_Z))
#define OUT_THREADS 128
#ifndef AWE_OUT_MINB
#define AWE_OUT_MINB 2          /* resident CTAs per SM the kernel is compiled for (register budget) */
#endif

__global__ void __launch_bounds__(OUT_THREADS, AWE_OUT_MINB) awegpu_out_kernel(const OutArgs a)
{
    const int n_nodes = a.n_k * (N_D + 1);
    const long long t = (long long)blockIdx.x * OUT_THREADS + threadIdx.x;
    if (t >= (long long)a.B * n_nodes) return;
    const int b = (int)(t / n_nodes), n = (int)(t - (long long)b * n_nodes);
    const int k = n / (N_D + 1), jj = n - k * (N_D + 1) - 1;           // jj = -1: shooting node
    const double *__restrict__ Wb = a.W + (size_t)b * a.n_v;
    const double *__restrict__ th0 = a.P + (size_t)b * a.ld_p + a.p_off;
    const double *__restrict__ slab = Wb + a.v_x0 + (size_t)k * SLAB_W;
    // x, u, z are read where the program uses them (read-only loads: the compiler may repeat a load instead of keeping
    // the value in a register); xdot and theta are formed once
    const double *__restrict__ xp = jj < 0 ? slab : slab + OFF_CX(0) + (size_t)jj * (N_X + N_Z);
    const double *__restrict__ zp = jj < 0 ? slab + OFF_Z : xp + N_X;
    double xd[N_X], th[N_TH];
    if (jj < 0) {
#pragma unroll
        for (int m = 0; m < N_X; ++m) xd[m] = __ldg(slab + OFF_XDOT + m);
    } else {
        const double s = (double)a.n_k / __ldg(Wb + __ldg(a.tf_idx + k));
        double cc[N_D + 1];
#pragma unroll
        for (int r = 0; r <= N_D; ++r) cc[r] = s * __ldg(a.coeff + r * (N_D + 1) + jj + 1);
#pragma unroll
        for (int m = 0; m < N_X; ++m) {
            double acc = cc[0] * __ldg(slab + m);
#pragma unroll
            for (int r = 1; r <= N_D; ++r) acc = fma(cc[r], __ldg(slab + OFF_CX(r - 1) + m), acc);
            xd[m] = acc;
        }
    }
#pragma unroll
    for (int c = 0; c < N_TH; ++c) th[c] = __ldg(Wb + __ldg(a.theta_idx + k * N_TH + c));
    double *__restrict__ outp = a.Out + (size_t)b * AWE_N_OUT * n_nodes + n;
    const double *__restrict__ phi = Wb + a.v_phi;
#define VIN(i) ((i) < N_X ? __ldg(xp + (i)) : (i) < 2 * N_X ? xd[(i) - N_X] : (i) < 2 * N_X + N_U ? __ldg(slab + OFF_U + ((i) - 2 * N_X)) \
                : (i) < 2 * N_X + N_U + N_Z ? __ldg(zp + ((i) - 2 * N_X - N_U)) : th[(i) - 2 * N_X - N_U - N_Z])
#define PIN(i) ((i) < AWE_N_THETA0 ? __ldg(th0 + (i)) : __ldg(phi + ((i) - AWE_N_THETA0)))
#define OUT(i, val) outp[(size_t)(i) * n_nodes] = (val)
#include "node_out.inc"
#undef VIN
#undef PIN
#undef OUT
}

int awe_out_launch(cudaStream_t st, const OutArgs &a)
{
    const long long threads = (long long)a.B * a.n_k * (N_D + 1);
    if (threads <= 0) return 0;
    const long long blocks = (threads + OUT_THREADS - 1) / OUT_THREADS;
    awegpu_out_kernel<<<(unsigned)blocks, OUT_THREADS, 0, st>>>(a);
    return (int)cudaGetLastError();
}

int awe_out_n(void) { return AWE_N_OUT; }

const char *awe_out_label(int i) { return (i >= 0 && i < AWE_N_OUT) ? awe_out_labels[i] : nullptr; }
