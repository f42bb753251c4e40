// awegpu_kernels.cu — sm_100a kernels and C ABI (include/awegpu.h) of the collocation g / jac_g
// evaluator.  Compiled once per model spec: the node arithmetic comes from the generated
// node_coll.inc / node_shoot.inc (awebox_b200/codegen.py); everything else here is hand-written.
//
// Work decomposition:
//   grid  = (ceil(B / TILE), n_k)            one CTA = TILE OCP instances x one collocation interval k
//   block = 32 * (d + 1) threads             warp j < d : collocation node (k, j)  (lane = instance)
//                                            warp d     : shooting node k, continuity rows of k, constant
//                                                         Jacobian entries, periodicity rows (k = n_k-1)
// The interval's slice of V (x_k u_k xdot_k z_k {x,z}_{k,0..d-1} x_{k+1}, contiguous in V:
// awebox/ocp/var_struct.py:49-93) is staged into shared memory with coalesced loads, one padded row
// per instance (odd row stride -> conflict-free lane-strided FP64 reads), together with theta, theta0
// and phi.  Each lane then runs the straight-line FP64 node function on its instance and writes
// residual rows and Jacobian entries straight into their CCS slots through the per-(k,j) slot table.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>

#include "awegpu.h"
#include "model_defs.h"

#define TILE 32
#ifndef AWE_MIN_BLOCKS
#define AWE_MIN_BLOCKS 2
#endif
#define N_X AWE_N_X
#define N_U AWE_N_U
#define N_Z AWE_N_Z
#define N_TH AWE_N_THETA
#define N_D AWE_D
#define N_WARPS (AWE_D + 1)
#define N_THREADS (32 * N_WARPS)

// offsets inside the interval slab of V
#define OFF_X0 0
#define OFF_U (N_X)
#define OFF_XDOT (N_X + N_U)
#define OFF_Z (2 * N_X + N_U)
#define OFF_CX(j) (2 * N_X + N_U + N_Z + (j) * (N_X + N_Z))
#define OFF_CZ(j) (OFF_CX(j) + N_X)
#define OFF_XNEXT (2 * N_X + N_U + N_Z + N_D * (N_X + N_Z))
#define SLAB (OFF_XNEXT + N_X)
#define SLAB_STRIDE (SLAB | 1)
// per-instance globals: theta (N_TH), then [theta0; phi] (AWE_N_PARAM)
#define GLOB (N_TH + AWE_N_PARAM)
#define GLOB_STRIDE (GLOB | 1)
#define XD_STRIDE (N_X | 1)
#define SMEM_DOUBLES (TILE * SLAB_STRIDE + TILE * GLOB_STRIDE + N_D * TILE * XD_STRIDE)

struct DevProblem {
    int n_v, n_p, n_g, n_k, n_theta0, n_phi, v_phi, p_theta0;
    long long nnz;
    int n_coll_out, n_shoot_out, n_shoot_eq, n_ineq, n_periodic, row_periodic, x_last, x_first;
    const int *x_base, *theta_idx, *tf_idx, *row_shoot, *row_path, *row_cont, *row_coll, *periodic_idx;
    const int *coll_slot, *shoot_slot, *const_slot, *const_begin;
    const double *const_val;
    double cscale;
};

__constant__ double c_coeff_coll[(N_D + 1) * (N_D + 1)];
__constant__ double c_coeff_cont[N_D + 1];

#define RSQRT(x) rsqrt(x)
#define SIGN(x) ((double)(((x) > 0.0) - ((x) < 0.0)))
#define PIN(i) gl[N_TH + (i)]
#define CCIN(r) cc[r]
#define OWNIN(r) own[r]
#define ITF itf

// --------------------------------------------------------------------------- collocation node
__device__ __forceinline__ void node_coll(const double *__restrict__ xs, const double *__restrict__ xd,
                                          const double *__restrict__ us, const double *__restrict__ zs,
                                          const double *__restrict__ gl, const double (&cc)[N_D + 1],
                                          const double (&own)[N_D + 1], const double itf,
                                          const int *__restrict__ tab, double *__restrict__ gout, size_t gstride,
                                          double *__restrict__ jout, size_t jstride)
{
#define VIN(i) ((i) < N_X ? xs[(i)] : (i) < 2 * N_X ? xd[(i) - N_X] : (i) < 2 * N_X + N_U ? us[(i) - 2 * N_X] \
                : (i) < 2 * N_X + N_U + N_Z ? zs[(i) - 2 * N_X - N_U] : gl[(i) - 2 * N_X - N_U - N_Z])
#define OUT(n, val)                                                        \
    do {                                                                   \
        if ((n) < AWE_N_EQ) {                                              \
            gout[(size_t)(n) * gstride] = (val);                           \
        } else if (jout) {                                                 \
            const int sl_ = __ldg(tab + (n));                              \
            if (sl_ >= 0) jout[(size_t)sl_ * jstride] = (val);             \
        }                                                                  \
    } while (0)
#include "node_coll.inc"
#undef OUT
}

// --------------------------------------------------------------------------- shooting node
__device__ __forceinline__ void node_shoot(const double *__restrict__ xs, const double *__restrict__ xd,
                                           const double *__restrict__ us, const double *__restrict__ zs,
                                           const double *__restrict__ gl, const int *__restrict__ tab,
                                           double *__restrict__ gout_eq, double *__restrict__ gout_ineq,
                                           size_t gstride, double *__restrict__ jout, size_t jstride, int n_shoot_eq)
{
#define OUT(n, val)                                                        \
    do {                                                                   \
        if ((n) < AWE_N_SHOOT_ROWS) {                                      \
            if ((n) < n_shoot_eq) gout_eq[(size_t)(n) * gstride] = (val);  \
            else gout_ineq[(size_t)((n) - n_shoot_eq) * gstride] = (val);  \
        } else if (jout) {                                                 \
            const int sl_ = __ldg(tab + (n));                              \
            if (sl_ >= 0) jout[(size_t)sl_ * jstride] = (val);             \
        }                                                                  \
    } while (0)
#include "node_shoot.inc"
#undef OUT
#undef VIN
}

__global__ void __launch_bounds__(N_THREADS, AWE_MIN_BLOCKS)
awegpu_eval_kernel(const DevProblem pr, const int B, const double *__restrict__ W, const double *__restrict__ P,
                   const long long ldp, const int p_off, double *__restrict__ G, double *__restrict__ Jv,
                   const int batch_minor)
{
    extern __shared__ double smem[];
    double *slab = smem;
    double *glob = smem + TILE * SLAB_STRIDE;
    double *xdot = glob + TILE * GLOB_STRIDE;

    const int k = blockIdx.y;
    const int b0 = blockIdx.x * TILE;
    const int tid = threadIdx.x;
    const int x0 = pr.x_base[k * (N_D + 1)];

    // ---- stage the interval slab of every instance of the tile (coalesced along the slab)
    for (int e = tid; e < TILE * SLAB; e += N_THREADS) {
        const int u = e / SLAB, o = e - u * SLAB;
        const int b = b0 + u;
        if (b < B) slab[u * SLAB_STRIDE + o] = __ldg(W + (size_t)b * pr.n_v + x0 + o);
    }
    for (int e = tid; e < TILE * GLOB; e += N_THREADS) {
        const int u = e / GLOB, o = e - u * GLOB;
        const int b = b0 + u;
        if (b < B) {
            double val;
            if (o < N_TH)
                val = __ldg(W + (size_t)b * pr.n_v + pr.theta_idx[k * N_TH + o]);
            else if (o < N_TH + pr.n_theta0)
                val = __ldg(P + (size_t)b * ldp + p_off + (o - N_TH));
            else
                val = __ldg(W + (size_t)b * pr.n_v + pr.v_phi + (o - N_TH - pr.n_theta0));
            glob[u * GLOB_STRIDE + o] = val;
        }
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int b = b0 + lane;
    if (b >= B) return;
    const double *s = slab + lane * SLAB_STRIDE;
    const double *gl = glob + lane * GLOB_STRIDE;
    const size_t gstride = batch_minor ? (size_t)B : 1;
    double *gbase = batch_minor ? (G + b) : (G + (size_t)b * pr.n_g);
    double *jbase = Jv ? (batch_minor ? (Jv + b) : (Jv + (size_t)b * pr.nnz)) : nullptr;
    const double tf = gl[AWE_TF_LOCAL - (2 * N_X + N_U + N_Z)];

    if (warp < N_D) {
        const int j = warp;
        double *xd = xdot + (j * TILE + lane) * XD_STRIDE;
        double cc[N_D + 1], own[N_D + 1];
        const double inv_h = (double)pr.n_k;
        const double itf = 1.0 / tf;
#pragma unroll
        for (int r = 0; r <= N_D; ++r) {
            cc[r] = c_coeff_coll[r * (N_D + 1) + (j + 1)] * inv_h * itf;
            own[r] = (r == j + 1) ? 1.0 : 0.0;
        }
        // xdot_kj = (sum_r C[r, j+1] X_r) / h / t_f      (awebox/ocp/collocation.py:225, 239-250)
        for (int m = 0; m < N_X; ++m) {
            double xp = c_coeff_coll[0 * (N_D + 1) + (j + 1)] * s[OFF_X0 + m];
#pragma unroll
            for (int r = 1; r <= N_D; ++r) xp += c_coeff_coll[r * (N_D + 1) + (j + 1)] * s[OFF_CX(r - 1) + m];
            xd[m] = xp * inv_h * itf;
        }
        const int *tab = pr.coll_slot + (size_t)(k * N_D + j) * pr.n_coll_out;
        node_coll(s + OFF_CX(j), xd, s + OFF_U, s + OFF_CZ(j), gl, cc, own, itf, tab,
                  gbase + (size_t)pr.row_coll[k * N_D + j] * gstride, gstride, jbase, gstride == 1 ? 1 : (size_t)B);
    } else {
        const int *tab = pr.shoot_slot + (size_t)k * pr.n_shoot_out;
        const size_t jstride = batch_minor ? (size_t)B : 1;
        node_shoot(s + OFF_X0, s + OFF_XDOT, s + OFF_U, s + OFF_Z, gl, tab,
                   gbase + (size_t)pr.row_shoot[k] * gstride, gbase + (size_t)pr.row_path[k] * gstride, gstride,
                   jbase, jstride, pr.n_shoot_eq);
        // continuity rows: x_{k+1} - sum_r D[r] X_r        (awebox/ocp/collocation.py:319-336)
        double *gc = gbase + (size_t)pr.row_cont[k] * gstride;
        for (int m = 0; m < N_X; ++m) {
            double xf = c_coeff_cont[0] * s[OFF_X0 + m];
#pragma unroll
            for (int r = 1; r <= N_D; ++r) xf += c_coeff_cont[r] * s[OFF_CX(r - 1) + m];
            gc[(size_t)m * gstride] = pr.cscale * (s[OFF_XNEXT + m] - xf);
        }
        // constant Jacobian entries owned by this interval
        if (jbase) {
            const int e1 = pr.const_begin[k + 1];
            for (int e = pr.const_begin[k]; e < e1; ++e)
                jbase[(size_t)__ldg(pr.const_slot + e) * jstride] = __ldg(pr.const_val + e);
        }
        // periodicity rows  x_0 - x_{n_k-1, d-1}           (awebox/ocp/operation.py:245-267)
        if (k == pr.n_k - 1) {
            const double *w = W + (size_t)b * pr.n_v;
            double *gp = gbase + (size_t)pr.row_periodic * gstride;
            for (int n = 0; n < pr.n_periodic; ++n) {
                const int m = __ldg(pr.periodic_idx + n);
                gp[(size_t)n * gstride] = pr.cscale * (__ldg(w + pr.x_first + m) - s[OFF_CX(N_D - 1) + m]);
            }
        }
    }
}

// =========================================================================== host side / C ABI
struct awegpu_handle {
    int device = 0;
    DevProblem dp{};
    awegpu_problem_t host{};
    std::vector<void *> dev_allocs;
    std::vector<long long> colind, row;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = -1.f;
    long long launches = 0;
    std::string err;
    // staging for the host entry
    double *dW = nullptr, *dP = nullptr, *dG = nullptr, *dJ = nullptr;
    double *hW = nullptr, *hP = nullptr, *hG = nullptr, *hJ = nullptr;
    size_t capW = 0, capP = 0, capG = 0, capJ = 0, hcapW = 0, hcapP = 0, hcapG = 0, hcapJ = 0;
    cudaStream_t stream2 = nullptr;
    // casadi sparsity buffers
    std::vector<long long> sp_v, sp_p, sp_g, sp_jv, sp_jp;
    std::vector<double> scratch_g;
};

static std::string g_create_error;
static awegpu_handle *g_casadi = nullptr;

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            set_err(h, std::string(#call) + ": " + cudaGetErrorString(e_));              \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

static void set_err(awegpu_handle *h, const std::string &m)
{
    if (h) h->err = m;
    g_create_error = m;
}

template <typename T>
static int to_device(awegpu_handle *h, const T *src, size_t n, const T **dst)
{
    void *p = nullptr;
    CK(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
    h->dev_allocs.push_back(p);
    if (n) CK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T *)p;
    return 0;
}

extern "C" const char *awegpu_model_name(void) { return AWE_MODEL_NAME; }

extern "C" int awegpu_model_dims(int32_t *n_var, int32_t *n_param, int32_t *n_coll_out, int32_t *n_shoot_out,
                                 int64_t *flops_coll, int64_t *flops_shoot)
{
    if (n_var) *n_var = AWE_N_VAR;
    if (n_param) *n_param = AWE_N_PARAM;
    if (n_coll_out) *n_coll_out = AWE_N_COLL_OUT;
    if (n_shoot_out) *n_shoot_out = AWE_N_SHOOT_OUT;
    if (flops_coll) *flops_coll = AWE_FLOPS_COLL;
    if (flops_shoot) *flops_shoot = AWE_FLOPS_SHOOT;
    return 0;
}

extern "C" const char *awegpu_last_error(awegpu_handle_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" void awegpu_destroy(awegpu_handle_t *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (g_casadi == h) g_casadi = nullptr;
    for (void *p : h->dev_allocs) cudaFree(p);
    cudaFree(h->dW), cudaFree(h->dP), cudaFree(h->dG), cudaFree(h->dJ);
    cudaFreeHost(h->hW), cudaFreeHost(h->hP), cudaFreeHost(h->hG), cudaFreeHost(h->hJ);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    delete h;
}

static int create_impl(awegpu_handle *h, const awegpu_problem_t *p, int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_err(h, std::string("no usable CUDA device (") + cudaGetErrorString(e) + "): awegpu has no CPU fallback");
        return 3;
    }
    if (device < 0 || device >= count) {
        set_err(h, "device index out of range");
        return 3;
    }
    h->device = device;
    CK(cudaSetDevice(device));
    // the compiled node functions must belong to this problem
    if (p->n_var != AWE_N_VAR || p->n_theta0 != AWE_N_THETA0 || p->n_phi != AWE_N_PHI || p->d != AWE_D ||
        p->n_x != N_X || p->n_u != N_U || p->n_z != N_Z || p->n_theta != N_TH ||
        p->n_coll_out != AWE_N_COLL_OUT || p->n_shoot_out != AWE_N_SHOOT_OUT ||
        p->n_shoot_eq + p->n_ineq != AWE_N_SHOOT_ROWS) {
        set_err(h, "problem dimensions do not match the compiled model '" AWE_MODEL_NAME "'");
        return 4;
    }
    // the kernel relies on the contiguous interval layout of V (var_struct.py:49-93)
    for (int k = 0; k < p->n_k; ++k) {
        const int32_t *xb = p->x_base + k * (p->d + 1);
        bool ok = p->u_base[k] == xb[0] + OFF_U && p->xdot_base[k] == xb[0] + OFF_XDOT && p->z_base[k] == xb[0] + OFF_Z &&
                  p->x_next[k] == xb[0] + OFF_XNEXT;
        for (int j = 0; j < p->d; ++j)
            ok = ok && xb[j + 1] == xb[0] + OFF_CX(j) && p->cz_base[k * p->d + j] == xb[0] + OFF_CZ(j);
        if (!ok) {
            set_err(h, "V layout is not the contiguous interval layout this kernel was written for");
            return 4;
        }
    }
    h->host = *p;
    DevProblem &d = h->dp;
    d.n_v = p->n_v, d.n_p = p->n_p, d.n_g = p->n_g, d.n_k = p->n_k, d.n_theta0 = p->n_theta0, d.n_phi = p->n_phi;
    d.v_phi = p->v_phi, d.p_theta0 = p->p_theta0, d.nnz = p->nnz;
    d.n_coll_out = p->n_coll_out, d.n_shoot_out = p->n_shoot_out, d.n_shoot_eq = p->n_shoot_eq, d.n_ineq = p->n_ineq;
    d.n_periodic = p->n_periodic, d.row_periodic = p->row_periodic, d.x_last = p->x_last, d.x_first = p->x_base[0];
    d.cscale = p->constraint_scale;
    const int nk = p->n_k, dd = p->d;
    int rc = 0;
    rc |= to_device(h, p->x_base, (size_t)nk * (dd + 1), &d.x_base);
    rc |= to_device(h, p->theta_idx, (size_t)nk * p->n_theta, &d.theta_idx);
    rc |= to_device(h, p->tf_idx, (size_t)nk, &d.tf_idx);
    rc |= to_device(h, p->row_shoot, (size_t)nk, &d.row_shoot);
    rc |= to_device(h, p->row_path, (size_t)nk, &d.row_path);
    rc |= to_device(h, p->row_cont, (size_t)nk, &d.row_cont);
    rc |= to_device(h, p->row_coll, (size_t)nk * dd, &d.row_coll);
    rc |= to_device(h, p->periodic_idx, (size_t)p->n_periodic, &d.periodic_idx);
    rc |= to_device(h, p->coll_slot, (size_t)nk * dd * p->n_coll_out, &d.coll_slot);
    rc |= to_device(h, p->shoot_slot, (size_t)nk * p->n_shoot_out, &d.shoot_slot);
    rc |= to_device(h, p->const_slot, (size_t)p->n_const, &d.const_slot);
    rc |= to_device(h, p->const_begin, (size_t)nk + 1, &d.const_begin);
    rc |= to_device(h, p->const_val, (size_t)p->n_const, &d.const_val);
    if (rc) return rc;
    CK(cudaMemcpyToSymbol(c_coeff_coll, p->coeff_coll, sizeof(double) * (dd + 1) * (dd + 1)));
    CK(cudaMemcpyToSymbol(c_coeff_cont, p->coeff_cont, sizeof(double) * (dd + 1)));
    h->colind.assign(p->colind, p->colind + p->n_v + 1);
    h->row.assign(p->row, p->row + p->nnz);
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    CK(cudaFuncSetAttribute(awegpu_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(SMEM_DOUBLES * sizeof(double))));
    // host copies of pointers are not kept: null them so nobody dereferences caller memory later
    h->host.x_base = h->host.x_next = h->host.u_base = h->host.xdot_base = h->host.z_base = nullptr;
    return 0;
}

extern "C" int awegpu_create(const awegpu_problem_t *problem, int device, awegpu_handle_t **out)
{
    if (!problem || !out) return 1;
    awegpu_handle *h = new awegpu_handle();
    int rc = create_impl(h, problem, device);
    if (rc) {
        g_create_error = h->err;
        awegpu_destroy(h);
        *out = nullptr;
        return rc;
    }
    *out = h;
    return 0;
}

extern "C" int awegpu_sparsity(awegpu_handle_t *h, int64_t *colind, int64_t *row)
{
    if (!h) return 1;
    if (colind) memcpy(colind, h->colind.data(), h->colind.size() * sizeof(int64_t));
    if (row) memcpy(row, h->row.data(), h->row.size() * sizeof(int64_t));
    return 0;
}

extern "C" int awegpu_eval_batch(awegpu_handle_t *h, int32_t B, const double *W, const double *P, double *G,
                                 double *Jv, int32_t flags, void *stream)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    if (!W || !P || !G) {
        set_err(h, "null W / P / G");
        return 1;
    }
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    long long ldp;
    int p_off;
    if (flags & AWEGPU_P_THETA0) {
        ldp = (flags & AWEGPU_P_SHARED) ? 0 : h->dp.n_theta0;
        p_off = 0;
    } else {
        ldp = (flags & AWEGPU_P_SHARED) ? 0 : h->dp.n_p;
        p_off = h->dp.p_theta0;
    }
    double *jv = (flags & AWEGPU_SKIP_JAC) ? nullptr : Jv;
    dim3 grid((B + TILE - 1) / TILE, h->dp.n_k);
    CK(cudaEventRecord(h->ev0, st));
    awegpu_eval_kernel<<<grid, N_THREADS, SMEM_DOUBLES * sizeof(double), st>>>(
        h->dp, B, W, P, ldp, p_off, G, jv, (flags & AWEGPU_JAC_BATCH_MINOR) ? 1 : 0);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, st));
    h->launches += 1;
    h->last_ms = -2.f;
    return 0;
}

extern "C" int awegpu_synchronize(awegpu_handle_t *h)
{
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" float awegpu_last_eval_ms(awegpu_handle_t *h)
{
    if (!h || h->last_ms == -1.f) return -1.f;
    if (h->last_ms == -2.f) {
        if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.f;
        float ms = -1.f;
        if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.f;
        h->last_ms = ms;
    }
    return h->last_ms;
}

extern "C" int64_t awegpu_launch_count(awegpu_handle_t *h) { return h ? h->launches : 0; }

static int grow(awegpu_handle *h, double **dev, double **host, size_t *cap, size_t *hcap, size_t need, bool with_host)
{
    if (need > *cap) {
        if (*dev) cudaFree(*dev);
        *dev = nullptr, *cap = 0;
        CK(cudaMalloc((void **)dev, need * sizeof(double)));
        *cap = need;
    }
    if (with_host && need > *hcap) {
        if (*host) cudaFreeHost(*host);
        *host = nullptr, *hcap = 0;
        CK(cudaMallocHost((void **)host, need * sizeof(double)));
        *hcap = need;
    }
    return 0;
}

extern "C" int awegpu_eval_batch_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P, double *G,
                                      double *Jv, int32_t flags)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    CK(cudaSetDevice(h->device));
    const DevProblem &d = h->dp;
    const bool want_j = Jv && !(flags & AWEGPU_SKIP_JAC);
    const size_t nW = (size_t)B * d.n_v;
    const size_t p_len = (flags & AWEGPU_P_THETA0) ? d.n_theta0 : d.n_p;
    const size_t nP = (flags & AWEGPU_P_SHARED) ? p_len : (size_t)B * p_len;
    const size_t nG = (size_t)B * d.n_g, nJ = want_j ? (size_t)B * d.nnz : 0;
    const bool stage = !(flags & AWEGPU_HOST_PINNED);
    int rc = grow(h, &h->dW, &h->hW, &h->capW, &h->hcapW, nW, stage) | grow(h, &h->dP, &h->hP, &h->capP, &h->hcapP, nP, stage) |
             grow(h, &h->dG, &h->hG, &h->capG, &h->hcapG, nG, stage) |
             (want_j ? grow(h, &h->dJ, &h->hJ, &h->capJ, &h->hcapJ, nJ, stage) : 0);
    if (rc) return rc;
    if (flags & AWEGPU_HOST_PINNED) {
        // caller buffers are page-locked: copy straight from / into them, in chunks on two streams so
        // that H2D of chunk c+1, the kernel of chunk c and D2H of chunk c-1 overlap.
        if (!h->stream2) CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
        const bool shared = (flags & AWEGPU_P_SHARED) != 0;
        if (shared) CK(cudaMemcpyAsync(h->dP, P, nP * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        const int chunk = 512;
        int c = 0;
        for (int b0 = 0; b0 < B; b0 += chunk, ++c) {
            const int nb = (B - b0 < chunk) ? (B - b0) : chunk;
            cudaStream_t st = (c & 1) ? h->stream2 : h->stream;
            CK(cudaMemcpyAsync(h->dW + (size_t)b0 * d.n_v, W + (size_t)b0 * d.n_v, (size_t)nb * d.n_v * sizeof(double),
                               cudaMemcpyHostToDevice, st));
            if (!shared)
                CK(cudaMemcpyAsync(h->dP + (size_t)b0 * p_len, P + (size_t)b0 * p_len, (size_t)nb * p_len * sizeof(double),
                                   cudaMemcpyHostToDevice, st));
            if (flags & AWEGPU_JAC_BATCH_MINOR) {
                set_err(h, "AWEGPU_HOST_PINNED supports the instance-major layout only");
                return 1;
            }
            rc = awegpu_eval_batch(h, nb, h->dW + (size_t)b0 * d.n_v, shared ? h->dP : h->dP + (size_t)b0 * p_len,
                                   h->dG + (size_t)b0 * d.n_g, want_j ? h->dJ + (size_t)b0 * d.nnz : nullptr, flags, st);
            if (rc) return rc;
            CK(cudaMemcpyAsync(G + (size_t)b0 * d.n_g, h->dG + (size_t)b0 * d.n_g, (size_t)nb * d.n_g * sizeof(double),
                               cudaMemcpyDeviceToHost, st));
            if (want_j)
                CK(cudaMemcpyAsync(Jv + (size_t)b0 * d.nnz, h->dJ + (size_t)b0 * d.nnz, (size_t)nb * d.nnz * sizeof(double),
                                   cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaStreamSynchronize(h->stream2));
        return 0;
    }
    memcpy(h->hW, W, nW * sizeof(double));
    memcpy(h->hP, P, nP * sizeof(double));
    CK(cudaMemcpyAsync(h->dW, h->hW, nW * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->dP, h->hP, nP * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rc = awegpu_eval_batch(h, B, h->dW, h->dP, h->dG, want_j ? h->dJ : nullptr, flags, h->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->hG, h->dG, nG * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (want_j) CK(cudaMemcpyAsync(h->hJ, h->dJ, nJ * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(G, h->hG, nG * sizeof(double));
    if (want_j) memcpy(Jv, h->hJ, nJ * sizeof(double));
    return 0;
}

// --------------------------------------------------------------------------- CasADi external ABI
static void dense_sp(std::vector<long long> &v, long long n)
{
    v = {n, 1, 1};     // compact dense flag: [nrow, ncol, 1]
}

extern "C" int awegpu_casadi_bind(awegpu_handle_t *h)
{
    if (!h) return 1;
    g_casadi = h;
    dense_sp(h->sp_v, h->dp.n_v);
    dense_sp(h->sp_p, h->dp.n_p);
    dense_sp(h->sp_g, h->dp.n_g);
    h->sp_jv.clear();
    h->sp_jv.push_back(h->dp.n_g);
    h->sp_jv.push_back(h->dp.n_v);
    h->sp_jv.insert(h->sp_jv.end(), h->colind.begin(), h->colind.end());
    h->sp_jv.insert(h->sp_jv.end(), h->row.begin(), h->row.end());
    h->sp_jp.assign(2 + (size_t)h->dp.n_p + 1, 0);      // n_g x n_p, no non-zeros
    h->sp_jp[0] = h->dp.n_g;
    h->sp_jp[1] = h->dp.n_p;
    return 0;
}

static int casadi_eval(const double **arg, double **res, bool jac)
{
    awegpu_handle *h = g_casadi;
    if (!h) return 1;
    if (!arg || !arg[0] || !arg[1]) {
        set_err(h, "awegpu_g: V and P must be given (NULL means all-zero, for which g is undefined)");
        return 1;
    }
    double *out = res ? res[0] : nullptr;
    if (!out) return 0;
    if (jac) h->scratch_g.resize(h->dp.n_g);
    return awegpu_eval_batch_host(h, 1, arg[0], arg[1], jac ? h->scratch_g.data() : out, jac ? out : nullptr,
                                  jac ? 0 : AWEGPU_SKIP_JAC);
}

extern "C" int awegpu_g(const double **arg, double **res, long long *, double *, int) { return casadi_eval(arg, res, false); }
extern "C" long long awegpu_g_n_in(void) { return 2; }
extern "C" long long awegpu_g_n_out(void) { return 1; }
extern "C" const char *awegpu_g_name_in(long long i) { return i == 0 ? "V" : i == 1 ? "P" : nullptr; }
extern "C" const char *awegpu_g_name_out(long long i) { return i == 0 ? "g" : nullptr; }
extern "C" const long long *awegpu_g_sparsity_in(long long i)
{
    if (!g_casadi) return nullptr;
    return i == 0 ? g_casadi->sp_v.data() : i == 1 ? g_casadi->sp_p.data() : nullptr;
}
extern "C" const long long *awegpu_g_sparsity_out(long long i) { return (g_casadi && i == 0) ? g_casadi->sp_g.data() : nullptr; }
extern "C" int awegpu_g_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w)
{
    if (sz_arg) *sz_arg = 2;
    if (sz_res) *sz_res = 1;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}

extern "C" int jac_awegpu_g(const double **arg, double **res, long long *, double *, int) { return casadi_eval(arg, res, true); }
extern "C" long long jac_awegpu_g_n_in(void) { return 3; }
extern "C" long long jac_awegpu_g_n_out(void) { return 2; }
extern "C" const char *jac_awegpu_g_name_in(long long i) { return i == 0 ? "V" : i == 1 ? "P" : i == 2 ? "out_g" : nullptr; }
extern "C" const char *jac_awegpu_g_name_out(long long i) { return i == 0 ? "jac_g_V" : i == 1 ? "jac_g_P" : nullptr; }
extern "C" const long long *jac_awegpu_g_sparsity_in(long long i)
{
    if (!g_casadi) return nullptr;
    return i == 0 ? g_casadi->sp_v.data() : i == 1 ? g_casadi->sp_p.data() : i == 2 ? g_casadi->sp_g.data() : nullptr;
}
extern "C" const long long *jac_awegpu_g_sparsity_out(long long i)
{
    if (!g_casadi) return nullptr;
    return i == 0 ? g_casadi->sp_jv.data() : i == 1 ? g_casadi->sp_jp.data() : nullptr;
}
extern "C" int jac_awegpu_g_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w)
{
    if (sz_arg) *sz_arg = 3;
    if (sz_res) *sz_res = 2;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}
