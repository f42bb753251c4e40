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

__global__ void __launch_bounds__(OUT_THREADS) awegpu_out_kernel(const OutArgs a)
{
    const int n_nodes = a.n_k * (N_D + 1);
    const long long t = (long long)blockIdx.x * OUT_THREADS + threadIdx.x;
    if (t >= (long long)a.B * n_nodes) return;
    const int b = (int)(t / n_nodes), n = (int)(t - (long long)b * n_nodes);
    const int k = n / (N_D + 1), jj = n - k * (N_D + 1) - 1;           // jj = -1: shooting node
    const double *__restrict__ Wb = a.W + (size_t)b * a.n_v;
    const double *__restrict__ th0 = a.P + (size_t)b * a.ld_p + a.p_off;
    const double *__restrict__ slab = Wb + a.v_x0 + (size_t)k * SLAB_W;
    double v[AWE_N_VAR];
    if (jj < 0) {
#pragma unroll
        for (int m = 0; m < N_X; ++m) v[m] = slab[m], v[N_X + m] = slab[OFF_XDOT + m];
#pragma unroll
        for (int c = 0; c < N_Z; ++c) v[2 * N_X + N_U + c] = slab[OFF_Z + c];
    } else {
        const double s = (double)a.n_k / Wb[__ldg(a.tf_idx + k)];
        double cc[N_D + 1];
#pragma unroll
        for (int r = 0; r <= N_D; ++r) cc[r] = s * __ldg(a.coeff + r * (N_D + 1) + jj + 1);
        const double *__restrict__ xj = slab + OFF_CX(0) + (size_t)jj * (N_X + N_Z);
#pragma unroll
        for (int m = 0; m < N_X; ++m) {
            double xd = cc[0] * slab[m];
#pragma unroll
            for (int r = 1; r <= N_D; ++r) xd = fma(cc[r], slab[OFF_CX(r - 1) + m], xd);
            v[m] = xj[m];
            v[N_X + m] = xd;
        }
#pragma unroll
        for (int c = 0; c < N_Z; ++c) v[2 * N_X + N_U + c] = xj[N_X + c];
    }
#pragma unroll
    for (int c = 0; c < N_U; ++c) v[2 * N_X + c] = slab[OFF_U + c];
#pragma unroll
    for (int c = 0; c < N_TH; ++c) v[2 * N_X + N_U + N_Z + c] = Wb[__ldg(a.theta_idx + k * N_TH + c)];
    double *__restrict__ outp = a.Out + (size_t)b * AWE_N_OUT * n_nodes + n;
    const double *__restrict__ phi = Wb + a.v_phi;
#define VIN(i) v[i]
#define PIN(i) ((i) < AWE_N_THETA0 ? __ldg(th0 + (i)) : phi[(i) - AWE_N_THETA0])
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
