// awegpu_hess.cu — second-order hook: upper triangle of the Hessian of lam' g(w, p) with respect to
// w, CCS, for a batch of OCP instances.  Replaces the nlp_hess_l function CasADi derives inside cas.nlpsol when the
// reference solves with the exact Hessian (awebox/opts/default.py:323; Lagrangian: awebox/tools/debug_operations.py:567-596).
// Tables and decomposition: awebox_b200/ocp/hessian.py.
//
//   awegpu_hess_node_kernel   CTA = one interval of one tile (32 instances): the interval slab of W and the interval's
//                             rows of lam are staged in shared memory with coalesced reads; warp j < d evaluates the
//                             node Hessian of collocation node (k, j) over the extended node variables
//                             (x, xsum = sum_r C[r, j+1] X_r, u, z, theta, phi), warp d that of the shooting node
//                             (generated programs node_hess_coll.inc / node_hess_shoot.inc, lane = instance);
//                             every value leaves as one coalesced 256-byte store into scr [tile][n_hv][32].
//   awegpu_hess_asm_kernel    H = A hv with the constant sparse assembly matrix (chain through the collocation
//                             polynomial, sums over the nodes of an interval and, for the theta / phi block, over all
//                             nodes): warp = 32 consecutive CCS slots, lane = instance for the sums (coalesced reads of
//                             scr, fixed order: deterministic, no atomics), transposed through shared memory, lane =
//                             slot for the instance-major stores (whole 256-byte runs of a row of H).
#include <cuda_runtime.h>
#include <cstdint>

#include "awegpu_hess.h"
#include "model_defs.h"
#include "hess_defs.h"
#include "awegpu_arith.cuh"

#define N_X AWE_N_X
#define N_U AWE_N_U
#define N_Z AWE_N_Z
#define N_TH AWE_N_THETA
#define N_PHI AWE_N_PHI
#define N_D AWE_D
#define OFF_X0 0
#define OFF_U (N_X)
#define OFF_XDOT (N_X + N_U)
#define OFF_Z (2 * N_X + N_U)
#define OFF_CX(j) (2 * N_X + N_U + N_Z + (j) * (N_X + N_Z))
#define OFF_CZ(j) (OFF_CX(j) + N_X)
#define SLAB_W (2 * N_X + N_U + N_Z + N_D * (N_X + N_Z))
#define SLAB_S (SLAB_W | 1)                                   /* odd row stride: lane-strided reads are conflict-free */
#define N_LAM (AWE_H_NLAM_S + N_D * AWE_N_EQ)                   /* multipliers of one interval: shooting eq | path | d x coll */
#define LAM_S (N_LAM | 1)
#define XS_S (N_X | 1)
#define TP_S ((N_TH + N_PHI) | 1)
#define HESS_THREADS (32 * (N_D + 1))
#define NODE_VALS (N_D * AWE_H_NEC + AWE_H_NES)
#define HESS_SMEM_BYTES (8 * 32 * (SLAB_S + LAM_S + N_D * XS_S + TP_S))
#ifndef AWE_HESS_MINB
/* resident CTAs per SM the node kernel is compiled for: two if their staging fits the shared memory of an SM */
#define AWE_HESS_MINB ((2 * HESS_SMEM_BYTES + 2048 <= 227 * 1024) ? 2 : 1)
#endif

__constant__ double c_hess_coll[(N_D + 1) * (N_D + 1)];       // C[r, j] (awebox_b200/ocp/collocation.py)
__constant__ double c_hess_quad[N_D];                         // quadrature weights of the collocation nodes

static size_t hess_node_smem()
{
    return sizeof(double) * 32 * ((size_t)SLAB_S + LAM_S + (size_t)N_D * XS_S + TP_S);
}

__global__ void __launch_bounds__(HESS_THREADS, AWE_HESS_MINB) awegpu_hess_node_kernel(const HessArgs a, const int tile0)
{
    extern __shared__ double sm[];
    double *slab = sm;                                  // [32][SLAB_S]
    double *lams = slab + 32 * SLAB_S;                  // [32][LAM_S]
    double *xsum = lams + 32 * LAM_S;                   // [N_D][32][XS_S]
    double *thph = xsum + N_D * 32 * XS_S;              // [32][TP_S]  theta | phi
    const int tile = tile0 + blockIdx.x / a.n_k, k = blockIdx.x % a.n_k;
    const int ltile = blockIdx.x / a.n_k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = tile * 32;
    // ---- stage the interval slab, the multipliers and theta / phi of the 32 instances (rows past B: the last instance)
    for (int i = warp; i < 32; i += N_D + 1) {
        const int b = min(b0 + i, a.B - 1);
        const double *wrow = a.W + (size_t)b * a.n_v;
        const double *src = wrow + a.v_x0 + (size_t)k * SLAB_W;
        for (int e = lane; e < SLAB_W; e += 32) slab[i * SLAB_S + e] = __ldg(src + e);
        const double *lsrc = a.Lam + (size_t)b * a.n_g + a.row0 + (size_t)k * a.row_stride;
        for (int e = lane; e < N_LAM; e += 32) lams[i * LAM_S + e] = __ldg(lsrc + e);
        for (int e = lane; e < N_TH + N_PHI; e += 32)
            thph[i * TP_S + e] = __ldg(wrow + (e < N_TH ? a.theta_idx[k * N_TH + e] : a.v_phi + e - N_TH));
    }
    __syncthreads();
    const int b = min(b0 + lane, a.B - 1);
    const double *pth0 = a.P + (size_t)b * a.ld_p + a.p_off;
    const double *sl = slab + lane * SLAB_S;
    const double *tp = thph + lane * TP_S;
    const double hnk = (double)a.n_k;
    double *out = a.scr + ((size_t)ltile * a.n_hv + (size_t)k * NODE_VALS) * 32 + lane;
#if AWE_H_OBJ
    // objective part (awebox_b200/ocp/hessian.py: _objective_term): references, weights and cost factors from the full
    // parameter vector, or zeros (and lam_f = 0) when the caller asked for the constraint part only
    const bool use_obj = a.Sigma != nullptr;
    const double *pobj = use_obj ? a.P + (size_t)b * a.ld_p : a.zeros;
    const double sigf = use_obj ? __ldg(a.Sigma + b) : 0.0;
    double tper = 0.0, tref = 0.0;
    for (int i = 0; i < a.n_t; ++i) {
        tper = fma(a.t_coef[i], __ldg(a.W + (size_t)b * a.n_v + a.t_idx[i]), tper);
        tref = fma(a.t_coef[i], __ldg(pobj + a.t_idx[i]), tref);
    }
    const double *pref = pobj + a.v_x0 + (size_t)k * SLAB_W;
    const int *thi = a.theta_idx + k * N_TH;
#define SIGF sigf
#define TPER tper
#define TREF tref
#define PW(i) __ldg(pobj + a.n_v + (i))
#define CF(i) __ldg(pobj + a.n_v + a.n_var + (i))
#endif
#define HNK hnk
#define PIN(i) ((i) < AWE_N_THETA0 ? __ldg(pth0 + (i)) : tp[N_TH + (i) - AWE_N_THETA0])
#define HOUT(n, v) out[(n) * 32] = (v)
    if (warp < N_D) {
        const int j = warp;
        double *xs = xsum + (j * 32 + lane) * XS_S;
#pragma unroll 1
        for (int m = 0; m < N_X; ++m) {
            double s = c_hess_coll[0 * (N_D + 1) + j + 1] * sl[OFF_X0 + m];
#pragma unroll
            for (int r = 1; r <= N_D; ++r) s = fma(c_hess_coll[r * (N_D + 1) + j + 1], sl[OFF_CX(r - 1) + m], s);
            xs[m] = s;
        }
        const double *xo = sl + OFF_CX(j), *zo = sl + OFF_CZ(j);
        const double *lm = lams + lane * LAM_S + AWE_H_NLAM_S + j * AWE_N_EQ;
        out += (size_t)j * AWE_H_NEC * 32;
#define VIN(i) ((i) < N_X ? xo[(i)] : (i) < 2 * N_X ? 0.0 : (i) < 2 * N_X + N_U ? sl[OFF_U + (i) - 2 * N_X] : \
                (i) < 2 * N_X + N_U + N_Z ? zo[(i) - 2 * N_X - N_U] : tp[(i) - 2 * N_X - N_U - N_Z])
#define XSIN(i) xs[(i)]
#define LIN(i) lm[(i)]
#if AWE_H_OBJ
        const double qj = c_hess_quad[j], w0 = (k == 0 && j == 0) ? 1.0 : 0.0;
        const double *rx = pref + OFF_CX(j), *rz = pref + OFF_CZ(j);
#define QJ qj
#define W0 w0
#define REF(i) ((i) < N_X ? __ldg(rx + (i)) : (i) < 2 * N_X ? 0.0 : (i) < 2 * N_X + N_U ? __ldg(pref + OFF_U + (i) - 2 * N_X) : \
                (i) < 2 * N_X + N_U + N_Z ? __ldg(rz + (i) - 2 * N_X - N_U) : __ldg(pobj + thi[(i) - 2 * N_X - N_U - N_Z]))
#endif
        {
// re-convergence point of the collocation warps of this CTA (generated only with HESS_SYNC_EVERY > 0)
#define HSYNC asm volatile("bar.sync 1, %0;" ::"n"(32 * N_D) : "memory");
#include "node_hess_coll.inc"
#undef HSYNC
        }
#undef VIN
#undef XSIN
#undef LIN
    } else {
        const double *lm = lams + lane * LAM_S;
        out += (size_t)N_D * AWE_H_NEC * 32;
#define VIN(i) ((i) < N_X ? sl[OFF_X0 + (i)] : (i) < 2 * N_X ? sl[OFF_XDOT + (i) - N_X] : (i) < 2 * N_X + N_U ? sl[OFF_U + (i) - 2 * N_X] : \
                (i) < 2 * N_X + N_U + N_Z ? sl[OFF_Z + (i) - 2 * N_X - N_U] : tp[(i) - 2 * N_X - N_U - N_Z])
#define LIN(i) lm[(i)]
        {
#define HSYNC
#include "node_hess_shoot.inc"
#undef HSYNC
        }
#undef VIN
#undef LIN
    }
#undef PIN
#undef HOUT
#undef HNK
}

#define ASM_WARPS 4
__global__ void __launch_bounds__(32 * ASM_WARPS) awegpu_hess_asm_kernel(const HessArgs a, const int tile0)
{
    __shared__ double stg[ASM_WARPS][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ltile = blockIdx.y, tile = tile0 + ltile;
    const long long s0 = ((long long)blockIdx.x * ASM_WARPS + warp) * 32;
    if (s0 >= a.nnz) return;
    const double *hv = a.scr + (size_t)ltile * a.n_hv * 32 + lane;
    const int n_s = (int)min((long long)32, a.nnz - s0);
    // the terms of the warp's 32 slots are one contiguous range of the tables: walk it four terms at a time (independent
    // loads in flight) and add every product to its slot's row (lane = instance; rows are private to the warp)
    const int p_lo = __ldg(a.asm_ptr + s0 + min(lane, n_s)), p_hi = __ldg(a.asm_ptr + s0 + min(lane + 1, n_s));
    const int T0 = __shfl_sync(0xffffffffu, p_lo, 0), T1 = __shfl_sync(0xffffffffu, p_hi, 31);
#pragma unroll
    for (int s = 0; s < 32; ++s) stg[warp][s][lane] = 0.0;
    int s = 0, s_end = __shfl_sync(0xffffffffu, p_hi, 0);       // terms below s_end belong to slot s
    for (int t = T0; t < T1; t += 4) {
        double c[4], v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int tt = min(t + u, T1 - 1);
            c[u] = (t + u < T1) ? __ldg(a.asm_coef + tt) : 0.0;
            v[u] = __ldg(hv + (size_t)__ldg(a.asm_idx + tt) * 32);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (t + u < T1) {
                while (t + u >= s_end) {
                    ++s;
                    s_end = __shfl_sync(0xffffffffu, p_hi, s);
                }
                stg[warp][s][lane] = fma(c[u], v[u], stg[warp][s][lane]);
            }
        }
    }
    __syncwarp();
    const int n_i = min(32, a.B - tile * 32);
    if (lane < n_s)
        for (int i = 0; i < n_i; ++i) a.H[(size_t)(tile * 32 + i) * a.nnz + s0 + lane] = stg[warp][lane][i];
}

size_t awe_hess_scratch_per_tile(int n_hv) { return sizeof(double) * 32 * (size_t)n_hv; }

int awe_hess_set_coeff(const double *C, const double *quad)
{
    cudaError_t e = cudaMemcpyToSymbol(c_hess_coll, C, sizeof(double) * (N_D + 1) * (N_D + 1));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_hess_quad, quad, sizeof(double) * N_D);
    return (int)e;
}

int awe_hess_has_objective(void) { return AWE_H_OBJ; }
void awe_hess_node_dims(int *n_ec, int *n_es) { *n_ec = AWE_H_NEC, *n_es = AWE_H_NES; }

int awe_hess_launch(cudaStream_t st, const HessArgs &a, int tile0, int n_tiles)
{
    static bool attr_set = false;
    const size_t smem = hess_node_smem();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(awegpu_hess_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    awegpu_hess_node_kernel<<<n_tiles * a.n_k, HESS_THREADS, smem, st>>>(a, tile0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const long long per_cta = 32LL * ASM_WARPS;
    dim3 grid((unsigned)((a.nnz + per_cta - 1) / per_cta), (unsigned)n_tiles);
    awegpu_hess_asm_kernel<<<grid, 32 * ASM_WARPS, 0, st>>>(a, tile0);
    return (int)cudaGetLastError();
}
