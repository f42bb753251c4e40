// awegpu_kernels.cu — sm_100a kernels and C ABI (include/awegpu.h) of the collocation g / jac_g
// evaluator.  Compiled once per model spec: the node arithmetic comes from the generated, already
// scheduled and register-allocated k_*.inc files (awebox_b200/codegen.py, awebox_b200/schedule.py);
// everything else here is hand-written.
//
// Work decomposition:
//   coll kernel   grid = (ceil(B/32), n_k), block = 32*d threads
//                 warp j = collocation node (k, j), lane = OCP instance of the 32-instance tile
//   shoot kernel  grid = (ceil(B/32), ceil(n_k/4)), block = 128 threads
//                 warp  = shooting node of interval k = 4*blockIdx.y + warp, lane = instance; also
//                 writes the continuity rows, the constant Jacobian entries of its interval and the
//                 periodicity rows (k = n_k-1)
// One CTA per SM, 255 registers per thread.  Each thread runs a straight-line FP64 program of ~8600
// operations whose working set (~420 FP64 values) does not fit the register file: the generator's
// Belady allocation keeps 100 values in registers and moves the rest through numbered scratch slots —
// the hot tier is a per-thread column of shared memory (slot-major, lane-minor: conflict-free), the
// cold tier goes to local memory.  Inputs are read where the schedule needs them straight from W
// (L1-cached sectors); results are staged 32 outputs x 32 instances at a time in shared memory and
// written instance-major with lane = output, so that entries that are adjacent in the CCS value array
// leave the SM in the same store instruction.
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
#define N_X AWE_N_X
#define N_U AWE_N_U
#define N_Z AWE_N_Z
#define N_TH AWE_N_THETA
#define N_D AWE_D
#define COLL_THREADS (32 * N_D)
#define SHOOT_WARPS 4
#define SHOOT_THREADS (32 * SHOOT_WARPS)

// offsets inside the interval slab of V (awebox/ocp/var_struct.py:49-93)
#define OFF_X0 0
#define OFF_U (N_X)
#define OFF_XDOT (N_X + N_U)
#define OFF_Z (2 * N_X + N_U)
#define OFF_CX(j) (2 * N_X + N_U + N_Z + (j) * (N_X + N_Z))
#define OFF_CZ(j) (OFF_CX(j) + N_X)
#define OFF_XNEXT (2 * N_X + N_U + N_Z + N_D * (N_X + N_Z))
#define OFF_XR(r) ((r) == 0 ? OFF_X0 : OFF_CX((r) - 1))

// per-thread constants kept in shared memory: cc[0..d], 1/t_f, theta[0..n_theta)
#define N_CST (N_D + 2 + N_TH)
// output staging: AWE_CHUNK outputs x 32 instances per warp, row stride odd (conflict-free both ways)
#define STG_STRIDE 33
#define STG_DOUBLES (AWE_CHUNK * STG_STRIDE)
#define CTAS_PER_SM 2

struct DevProblem {
    int n_v, n_p, n_g, n_k, n_theta0, n_phi, v_phi, p_theta0;
    long long nnz;
    int n_coll_out, n_shoot_out, n_shoot_eq, n_ineq, n_periodic, row_periodic, x_last, x_first;
    const int *x_base, *theta_idx, *tf_idx, *row_cont, *periodic_idx;
    const int *coll_dest, *shoot_dest, *const_slot, *const_begin;
    const double *const_val;
    double cscale;
};

__constant__ double c_coeff_coll[(N_D + 1) * (N_D + 1)];
__constant__ double c_coeff_cont[N_D + 1];

#define SIGN(x) ((double)(((x) > 0.0) - ((x) < 0.0)))
// Reciprocal / reciprocal square root from the MUFU seed (>= 20 bits) and two Newton steps, without
// the special-case paths of the CUDA library versions (their arguments here are norms, squared
// norms + eps, densities, times: finite, normal and positive).  Division and square root get one
// extra residual correction, which leaves them within 1 ulp of the correctly rounded result.
__device__ __forceinline__ double rcp_nr(const double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double div_nr(const double a, const double b)
{
    const double r = rcp_nr(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}
__device__ __forceinline__ double rsqrt_nr(const double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);
    y = fma(y * 0.5, e, y);
    e = fma(-x * y, y, 1.0);
    return fma(y * 0.5, e, y);
}
__device__ __forceinline__ double sqrt_nr(const double x)
{
    const double y = rsqrt_nr(x);
    const double s = x * y;
    return fma(fma(-s, s, x), 0.5 * y, s);
}
#define RSQRT(x) rsqrt_nr(x)
#define SQRTF(x) sqrt_nr(x)
#define DIVF(a, b) div_nr(a, b)

// ---- statement macros of the generated code -------------------------------------------------------
// inputs: asm volatile so that the compiler neither merges the generator's reloads into one
// long-lived register nor hoists them; the address is base register + immediate
#define LDG_(reg, ptr) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(reg) : "l"(ptr))
#define LDS_(reg, addr) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(reg) : "r"(addr))
#define STS_(addr, reg) asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(reg))
#define LD_XR(reg, r, m) LDG_(reg, w_slab + (OFF_XR(r) + (m)))
#define LD_XO(reg, m) LDG_(reg, w_own + (m))
#define LD_XD(reg, m) LDG_(reg, w_slab + (OFF_XDOT + (m)))
#define LD_U(reg, c) LDG_(reg, w_slab + (OFF_U + (c)))
#define LD_Z(reg, c) LDG_(reg, w_z + (c))
#define LD_P(reg, i) LDG_(reg, ((i) < AWE_N_THETA0 ? p_th0 + (i) : w_phi + ((i) - AWE_N_THETA0)))
#define LD_CC(reg, r) LDS_(reg, cst_addr + (r) * (8 * NTHR))
#define LD_ITF(reg) LDS_(reg, cst_addr + (N_D + 1) * (8 * NTHR))
#define LD_TH(reg, t) LDS_(reg, cst_addr + (N_D + 2 + (t)) * (8 * NTHR))
#define LD_OWN(reg, r) reg = (own_r == (r)) ? 1.0 : 0.0
#define SPILL_ST(slot, reg) STS_(scr_addr + (slot) * (8 * NTHR), reg)
#define SPILL_LD(reg, slot) LDS_(reg, scr_addr + (slot) * (8 * NTHR))
#define OVF_ST(slot, reg) ovf[slot] = reg
#define OVF_LD(reg, slot) reg = ovf[slot]
#define SYNC_HINT() __syncthreads()

// staged output: OUT parks the value of output n for this lane's instance, FLUSH writes one chunk of
// up to 32 outputs for the 32 instances of the tile with lane = output
#define OUT(n, val)                                                            \
    do {                                                                       \
        if (BATCH_MINOR) {                                                     \
            const int code_ = __ldg(dest + (n));                               \
            if (valid) {                                                       \
                if (code_ >= 0) {                                              \
                    if (Jv) Jv[(size_t)code_ * B + b] = (val);                 \
                } else if (code_ <= -2)                                        \
                    G[(size_t)(-2 - code_) * B + b] = (val);                   \
            }                                                                  \
        } else                                                                 \
            stg[((n) % AWE_CHUNK) * STG_STRIDE + lane] = (val);                       \
    } while (0)
#define FLUSH(c, cnt)                                                          \
    do {                                                                       \
        if (!BATCH_MINOR) flush_chunk(stg, dest + (c) * AWE_CHUNK, cnt, lane, n_valid, g_tile, j_tile, n_g, nnz); \
    } while (0)

// lane = (instance group, output): output o = lane % AWE_CHUNK of the chunk, instances i = lane / AWE_CHUNK,
// + 32 / AWE_CHUNK, ...: every store instruction writes AWE_CHUNK neighbouring outputs of 32 / AWE_CHUNK
// instances
__device__ __forceinline__ void flush_chunk(double *stg, const int *__restrict__ dest, const int cnt, const int lane,
                                            const int n_valid, double *__restrict__ g_tile, double *__restrict__ j_tile,
                                            const int n_g, const long long nnz)
{
    constexpr int GROUPS = 32 / AWE_CHUNK;
    __syncwarp();
    const int o = lane % AWE_CHUNK, i0 = lane / AWE_CHUNK;
    const int code = o < cnt ? __ldg(dest + o) : -1;
    double *p = nullptr;
    size_t stride = 0;
    if (code >= 0) {
        if (j_tile) p = j_tile + code, stride = (size_t)nnz;
    } else if (code <= -2) {
        p = g_tile + (-2 - code), stride = (size_t)n_g;
    }
    if (p) {
        const double *src = stg + o * STG_STRIDE;
#pragma unroll 4
        for (int i = i0; i < n_valid; i += GROUPS) p[(size_t)i * stride] = src[i];
    }
    __syncwarp();
}

// --------------------------------------------------------------------------- collocation nodes
template <bool FULL, bool BATCH_MINOR>
__global__ void __launch_bounds__(COLL_THREADS, CTAS_PER_SM)
awegpu_coll_kernel(const DevProblem pr, const int B, const double *__restrict__ W, const double *__restrict__ P,
                   const long long ldp, const int p_off, double *__restrict__ G, double *__restrict__ Jv)
{
    constexpr int NTHR = COLL_THREADS;
    constexpr int HOT = FULL ? AWE_K_COLL_HOT : AWE_K_COLL_RES_HOT;
    constexpr int COLD = FULL ? AWE_K_COLL_COLD : AWE_K_COLL_RES_COLD;
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, j = tid >> 5;
    const int k = blockIdx.y, b0 = blockIdx.x * TILE;
    const int b_raw = b0 + lane;
    const bool valid = b_raw < B;
    const int b = valid ? b_raw : B - 1;          // surplus lanes of a ragged tile recompute the last instance
    const int n_valid = min(TILE, B - b0);
    const int n_g = pr.n_g;
    const long long nnz = pr.nnz;
    double *stg = smem + (size_t)(HOT + N_CST) * NTHR + j * STG_DOUBLES;
    const unsigned scr_addr = (unsigned)__cvta_generic_to_shared(smem + tid);
    const unsigned cst_addr = (unsigned)__cvta_generic_to_shared(smem + (size_t)HOT * NTHR + tid);
    double ovf[COLD > 0 ? COLD : 1];

    const double *w = W + (size_t)b * pr.n_v;
    const double *w_slab = w + __ldg(pr.x_base + k * (N_D + 1));
    const double *w_own = w_slab + OFF_CX(j);
    const double *w_z = w_own + N_X;
    const double *w_phi = w + pr.v_phi;
    const double *p_th0 = P + (size_t)b * ldp + p_off;
    const int own_r = j + 1;
    {
        // per-lane constants: cc[r] = C[r, j+1] / h / t_f (awebox/ocp/collocation.py:225), 1/t_f, theta
        const double itf = 1.0 / __ldg(w + __ldg(pr.tf_idx + k));
        double *cst = smem + (size_t)HOT * NTHR + tid;
#pragma unroll
        for (int r = 0; r <= N_D; ++r) cst[r * NTHR] = c_coeff_coll[r * (N_D + 1) + (j + 1)] * (double)pr.n_k * itf;
        cst[(N_D + 1) * NTHR] = itf;
#pragma unroll
        for (int t = 0; t < N_TH; ++t) cst[(N_D + 2 + t) * NTHR] = __ldg(w + __ldg(pr.theta_idx + k * N_TH + t));
    }
    const int *dest = pr.coll_dest + (size_t)(k * N_D + j) * pr.n_coll_out;
    double *g_tile = G + (size_t)b0 * n_g;
    double *j_tile = Jv ? Jv + (size_t)b0 * nnz : nullptr;
    __syncwarp();
    if (FULL) {
#include "k_coll.inc"
    } else {
#include "k_coll_res.inc"
    }
}

// --------------------------------------------------------------------------- shooting nodes
template <bool FULL, bool BATCH_MINOR>
__global__ void __launch_bounds__(SHOOT_THREADS, CTAS_PER_SM)
awegpu_shoot_kernel(const DevProblem pr, const int B, const double *__restrict__ W, const double *__restrict__ P,
                    const long long ldp, const int p_off, double *__restrict__ G, double *__restrict__ Jv)
{
    constexpr int NTHR = SHOOT_THREADS;
    constexpr int HOT = FULL ? AWE_K_SHOOT_HOT : AWE_K_SHOOT_RES_HOT;
    constexpr int COLD = FULL ? AWE_K_SHOOT_COLD : AWE_K_SHOOT_RES_COLD;
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k_raw = blockIdx.y * SHOOT_WARPS + wid;
    const bool k_ok = k_raw < pr.n_k;
    const int k = k_ok ? k_raw : pr.n_k - 1;
    const int b0 = blockIdx.x * TILE;
    const int b_raw = b0 + lane;
    const bool valid = k_ok && b_raw < B;
    const int b = b_raw < B ? b_raw : B - 1;
    const int n_valid = k_ok ? min(TILE, B - b0) : 0;
    const int n_g = pr.n_g;
    const long long nnz = pr.nnz;
    double *stg = smem + (size_t)(HOT + N_CST) * NTHR + wid * STG_DOUBLES;
    const unsigned scr_addr = (unsigned)__cvta_generic_to_shared(smem + tid);
    const unsigned cst_addr = (unsigned)__cvta_generic_to_shared(smem + (size_t)HOT * NTHR + tid);
    double ovf[COLD > 0 ? COLD : 1];

    const double *w = W + (size_t)b * pr.n_v;
    const double *w_slab = w + __ldg(pr.x_base + k * (N_D + 1));
    const double *w_own = w_slab + OFF_X0;
    const double *w_z = w_slab + OFF_Z;
    const double *w_phi = w + pr.v_phi;
    const double *p_th0 = P + (size_t)b * ldp + p_off;
    {
        double *cst = smem + (size_t)HOT * NTHR + tid;
#pragma unroll
        for (int t = 0; t < N_TH; ++t) cst[(N_D + 2 + t) * NTHR] = __ldg(w + __ldg(pr.theta_idx + k * N_TH + t));
    }
    const int *dest = pr.shoot_dest + (size_t)k * pr.n_shoot_out;
    double *g_tile = G + (size_t)b0 * n_g;
    double *j_tile = Jv ? Jv + (size_t)b0 * nnz : nullptr;
    __syncwarp();
    {
        const int own_r = -1;
        (void)own_r;
        if (FULL) {
#include "k_shoot.inc"
        } else {
#include "k_shoot_res.inc"
        }
    }
    // ---- continuity rows  x_{k+1} - sum_r D[r] X_r   (awebox/ocp/collocation.py:319-336)
    const int row_cont = __ldg(pr.row_cont + k);
    if (BATCH_MINOR) {
        if (valid)
            for (int m = 0; m < N_X; ++m) {
                double xf = c_coeff_cont[0] * __ldg(w_slab + OFF_X0 + m);
#pragma unroll
                for (int r = 1; r <= N_D; ++r) xf += c_coeff_cont[r] * __ldg(w_slab + OFF_CX(r - 1) + m);
                G[(size_t)(row_cont + m) * B + b] = pr.cscale * (__ldg(w_slab + OFF_XNEXT + m) - xf);
            }
    } else {
        for (int m0 = 0; m0 < N_X; m0 += AWE_CHUNK) {
            const int cnt = min(AWE_CHUNK, N_X - m0);
            for (int m = m0; m < m0 + cnt; ++m) {
                double xf = c_coeff_cont[0] * __ldg(w_slab + OFF_X0 + m);
#pragma unroll
                for (int r = 1; r <= N_D; ++r) xf += c_coeff_cont[r] * __ldg(w_slab + OFF_CX(r - 1) + m);
                stg[(m - m0) * STG_STRIDE + lane] = pr.cscale * (__ldg(w_slab + OFF_XNEXT + m) - xf);
            }
            __syncwarp();
            {
                const int o = lane % AWE_CHUNK;
                if (o < cnt) {
                    double *p = g_tile + row_cont + m0 + o;
                    const double *src = stg + o * STG_STRIDE;
                    for (int i = lane / AWE_CHUNK; i < n_valid; i += 32 / AWE_CHUNK) p[(size_t)i * n_g] = src[i];
                }
            }
            __syncwarp();
        }
    }
    // ---- constant Jacobian entries owned by this interval (continuity; periodicity on the last one):
    //      the value does not depend on the instance, lane = entry
    if (Jv && k_ok) {
        const int e1 = __ldg(pr.const_begin + k + 1);
        for (int e = __ldg(pr.const_begin + k) + lane; e < e1; e += 32) {
            const int slot = __ldg(pr.const_slot + e);
            const double v = __ldg(pr.const_val + e);
            if (BATCH_MINOR) {
                for (int i = 0; i < n_valid; ++i) Jv[(size_t)slot * B + b0 + i] = v;
            } else {
                for (int i = 0; i < n_valid; ++i) j_tile[(size_t)i * nnz + slot] = v;
            }
        }
    }
    // ---- periodicity rows  x_0 - x_{n_k-1, d-1}           (awebox/ocp/operation.py:245-267)
    if (k_raw == pr.n_k - 1 && valid) {
        for (int n = 0; n < pr.n_periodic; ++n) {
            const int m = __ldg(pr.periodic_idx + n);
            const double v = pr.cscale * (__ldg(w + pr.x_first + m) - __ldg(w_slab + OFF_CX(N_D - 1) + m));
            if (BATCH_MINOR)
                G[(size_t)(pr.row_periodic + n) * B + b] = v;
            else
                G[(size_t)b * n_g + pr.row_periodic + n] = v;
        }
    }
}

template <bool FULL>
static constexpr size_t coll_smem_bytes()
{
    return sizeof(double) * ((size_t)((FULL ? AWE_K_COLL_HOT : AWE_K_COLL_RES_HOT) + N_CST) * COLL_THREADS + N_D * STG_DOUBLES);
}
template <bool FULL>
static constexpr size_t shoot_smem_bytes()
{
    return sizeof(double) *
           ((size_t)((FULL ? AWE_K_SHOOT_HOT : AWE_K_SHOOT_RES_HOT) + N_CST) * SHOOT_THREADS + SHOOT_WARPS * STG_DOUBLES);
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
    // destination codes of the node-local outputs: >= 0 slot of the CCS value array, -1 none,
    // <= -2 row (-2 - code) of g
    std::vector<int> coll_dest((size_t)nk * dd * p->n_coll_out), shoot_dest((size_t)nk * p->n_shoot_out);
    for (int kj = 0; kj < nk * dd; ++kj)
        for (int n = 0; n < p->n_coll_out; ++n)
            coll_dest[(size_t)kj * p->n_coll_out + n] =
                n < AWE_N_EQ ? -2 - (p->row_coll[kj] + n) : p->coll_slot[(size_t)kj * p->n_coll_out + n];
    for (int k = 0; k < nk; ++k)
        for (int n = 0; n < p->n_shoot_out; ++n)
            shoot_dest[(size_t)k * p->n_shoot_out + n] =
                n < p->n_shoot_eq                ? -2 - (p->row_shoot[k] + n)
                : n < p->n_shoot_eq + p->n_ineq ? -2 - (p->row_path[k] + n - p->n_shoot_eq)
                                                : p->shoot_slot[(size_t)k * p->n_shoot_out + n];
    int rc = 0;
    rc |= to_device(h, p->x_base, (size_t)nk * (dd + 1), &d.x_base);
    rc |= to_device(h, p->theta_idx, (size_t)nk * p->n_theta, &d.theta_idx);
    rc |= to_device(h, p->tf_idx, (size_t)nk, &d.tf_idx);
    rc |= to_device(h, p->row_cont, (size_t)nk, &d.row_cont);
    rc |= to_device(h, p->periodic_idx, (size_t)p->n_periodic, &d.periodic_idx);
    rc |= to_device(h, coll_dest.data(), coll_dest.size(), &d.coll_dest);
    rc |= to_device(h, shoot_dest.data(), shoot_dest.size(), &d.shoot_dest);
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
#define SET_SMEM(kern, bytes) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)))
    SET_SMEM((awegpu_coll_kernel<true, false>), coll_smem_bytes<true>());
    SET_SMEM((awegpu_coll_kernel<true, true>), coll_smem_bytes<true>());
    SET_SMEM((awegpu_coll_kernel<false, false>), coll_smem_bytes<false>());
    SET_SMEM((awegpu_coll_kernel<false, true>), coll_smem_bytes<false>());
    SET_SMEM((awegpu_shoot_kernel<true, false>), shoot_smem_bytes<true>());
    SET_SMEM((awegpu_shoot_kernel<true, true>), shoot_smem_bytes<true>());
    SET_SMEM((awegpu_shoot_kernel<false, false>), shoot_smem_bytes<false>());
    SET_SMEM((awegpu_shoot_kernel<false, true>), shoot_smem_bytes<false>());
#undef SET_SMEM
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
    const bool full = jv != nullptr, bm = (flags & AWEGPU_JAC_BATCH_MINOR) != 0;
    const int tiles = (B + TILE - 1) / TILE;
    dim3 gc(tiles, h->dp.n_k), gs(tiles, (h->dp.n_k + SHOOT_WARPS - 1) / SHOOT_WARPS);
    CK(cudaEventRecord(h->ev0, st));
#define LAUNCH(F, M)                                                                                               \
    do {                                                                                                           \
        awegpu_coll_kernel<F, M><<<gc, COLL_THREADS, coll_smem_bytes<F>(), st>>>(h->dp, B, W, P, ldp, p_off, G, jv);   \
        awegpu_shoot_kernel<F, M><<<gs, SHOOT_THREADS, shoot_smem_bytes<F>(), st>>>(h->dp, B, W, P, ldp, p_off, G, jv); \
    } while (0)
    if (full && !bm) LAUNCH(true, false);
    else if (full && bm) LAUNCH(true, true);
    else if (!full && !bm) LAUNCH(false, false);
    else LAUNCH(false, true);
#undef LAUNCH
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, st));
    h->launches += 2;
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
