// Microbenchmarks behind the roofline denominators and the design decisions of DESIGN.md (run on a B200 with
// scripts/ubench/run.sh; results are committed under profiles/):
//   fp64   vector-FP64 peak: independent DFMA chains, every SM full                       
//   write  HBM write bandwidth for the store patterns the output stage can use: plain 8-byte and 16-byte
//          coalesced stores, stores whose warp covers several short runs (partial sectors), half-written sectors,
//          and bulk asynchronous shared->global copies (cp.async.bulk) of 64 B .. 4 KB per operation
//   icache a straight-line FP64 body of 16 KB .. 256 KB of SASS run by 8 / 16 warps per SM in lock step or evenly
//          staggered: what a program larger than the 32 KB instruction cache costs at full grid
//   lat    dependent-issue latency of DFMA (one chain, one warp)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static float time_ms(void (*launch)(void *), void *arg, int reps)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch(arg);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch(arg);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

// ------------------------------------------------------------------------------------------------ fp64 peak
template <int CHAINS>
__global__ void __launch_bounds__(256) k_fp64(double *out, double x, double y, int iters)
{
    double a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
}
struct Fp64Arg { double *out; int ctas, iters, chains; };
static void launch_fp64(void *p)
{
    Fp64Arg *a = (Fp64Arg *)p;
    if (a->chains == 8) k_fp64<8><<<a->ctas, 256>>>(a->out, 1.0000001, 1e-9, a->iters);
    else if (a->chains == 1) k_fp64<1><<<a->ctas, 256>>>(a->out, 1.0000001, 1e-9, a->iters);
    else k_fp64<4><<<a->ctas, 256>>>(a->out, 1.0000001, 1e-9, a->iters);
}

__global__ void k_lat(double *out, long long *clk, double x, double y, int iters)
{
    double a = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 64; ++u) a = fma(a, x, y);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[0] = t1 - t0;
    if (a == 1.2345) out[0] = a;
}

// ------------------------------------------------------------------------------------------------ write bandwidth
// mode 0: 8 B per lane, warp writes 256 contiguous bytes; 1: 16 B per lane (512 B); 2: the warp's 32 entries fall into
// 5 runs of 6-7 entries that start at odd 8-byte offsets 24 entries apart (partial sectors at both ends of every run);
// 3: 16 B per lane, every other 16 B (half of every 32-byte sector is written); 4: as 0 with default (write-back) policy
__global__ void __launch_bounds__(256) k_write(double *buf, size_t n_doubles, int mode)
{
    const size_t warp = (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
    const size_t n_warps = (size_t)gridDim.x * (blockDim.x >> 5);
    const int lane = threadIdx.x & 31;
    if (mode == 0 || mode == 4) {
        for (size_t i = warp * 32 + lane; i < n_doubles; i += n_warps * 32) {
            if (mode == 0) __stcs(buf + i, 1.0); else buf[i] = 1.0;
        }
    } else if (mode == 1) {
        double2 *b2 = (double2 *)buf;
        for (size_t i = warp * 32 + lane; i < n_doubles / 2; i += n_warps * 32) __stcs(b2 + i, make_double2(1.0, 2.0));
    } else if (mode == 2) {
        // a block of 5 * 24 = 120 entries per warp iteration, of which 32 are written: run r covers entries
        // r*24 + 1 .. r*24 + 7 (6-7 entries)
        const int run = lane / 7 > 4 ? 4 : lane / 7, pos = lane - run * 7 + (run == 4 ? 0 : 0);
        const size_t off = (size_t)run * 24 + 1 + pos;
        for (size_t i = warp * 120; i + 120 <= n_doubles; i += n_warps * 120) __stcs(buf + i + off, 1.0);
    } else if (mode == 3) {
        double2 *b2 = (double2 *)buf;
        for (size_t i = warp * 32 + lane; i < n_doubles / 4; i += n_warps * 32) __stcs(b2 + 2 * i, make_double2(1.0, 2.0));
    }
}
struct WriteArg { double *buf; size_t n; int mode, ctas; };
static void launch_write(void *p)
{
    WriteArg *a = (WriteArg *)p;
    k_write<<<a->ctas, 256>>>(a->buf, a->n, a->mode);
}

// bulk asynchronous copies shared -> global: every lane of every warp issues one copy of `bytes` per iteration from a
// staging slice of its CTA; a CTA keeps at most 2 groups per thread in flight
__global__ void __launch_bounds__(128) k_bulk(char *buf, size_t total, int bytes)
{
    extern __shared__ __align__(128) char stage[];
    for (int i = threadIdx.x; i < 32768 / 8; i += blockDim.x) ((double *)stage)[i] = 1.0;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const size_t thread = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_threads = (size_t)gridDim.x * blockDim.x;
    const unsigned src = (unsigned)__cvta_generic_to_shared(stage) + (unsigned)((threadIdx.x * bytes) % (32768 - 4096));
    for (size_t o = thread * bytes; o + bytes <= total; o += n_threads * bytes) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(buf + o), "r"(src), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
struct BulkArg { char *buf; size_t total; int bytes, ctas; };
static void launch_bulk(void *p)
{
    BulkArg *a = (BulkArg *)p;
    k_bulk<<<a->ctas, 128, 32768>>>(a->buf, a->total, a->bytes);
}

// ------------------------------------------------------------------------------------------------ instruction cache
#define S4 a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
#define S16 S4 S4 S4 S4
#define S64 S16 S16 S16 S16
#define S256 S64 S64 S64 S64
#define S1024 S256 S256 S256 S256
#define S4096 S1024 S1024 S1024 S1024
#define S16384 S4096 S4096 S4096 S4096
#define ICACHE_KERNEL(NAME, BODY)                                                                         \
    __global__ void __launch_bounds__(512, 1) NAME(double *out, double x, double y, int reps, int delay)  \
    {                                                                                                     \
        double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;                                   \
        if (delay > 0) {                                                                                  \
            const long long t0 = clock64();                                                               \
            while (clock64() - t0 < (long long)(threadIdx.x >> 5) * delay) { }                            \
        }                                                                                                 \
        _Pragma("unroll 1") for (int r = 0; r < reps; ++r) { BODY }                                       \
        if (a0 + a1 + a2 + a3 == 1.2345) out[0] = a0;                                                     \
    }
ICACHE_KERNEL(k_ic_1k, S1024)
ICACHE_KERNEL(k_ic_4k, S4096)
ICACHE_KERNEL(k_ic_16k, S16384)
struct IcArg { double *out; int which, warps, reps, delay; };
static void launch_ic(void *p)
{
    IcArg *a = (IcArg *)p;
    if (a->which == 0) k_ic_1k<<<148, 32 * a->warps>>>(a->out, 1.0000001, 1e-9, a->reps, a->delay);
    else if (a->which == 1) k_ic_4k<<<148, 32 * a->warps>>>(a->out, 1.0000001, 1e-9, a->reps, a->delay);
    else k_ic_16k<<<148, 32 * a->warps>>>(a->out, 1.0000001, 1e-9, a->reps, a->delay);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, prop.multiProcessorCount, clk_khz);
    double *out;
    CK(cudaMalloc(&out, 1024));
    // ---- fp64
    for (int chains : {8, 4, 1}) {
        for (int cps : {2, 4, 8}) {
            Fp64Arg a{out, prop.multiProcessorCount * cps, 4096, chains};
            const float ms = time_ms(launch_fp64, &a, 5);
            const double fl = 2.0 * (double)a.ctas * 256 * a.iters * 16 * chains;
            printf("{\"bench\": \"fp64\", \"chains\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", chains, cps, ms, fl / ms * 1e-9);
        }
    }
    {
        long long *clk;
        CK(cudaMalloc(&clk, 8));
        k_lat<<<1, 32>>>(out, clk, 1.0000001, 1e-9, 1000);
        CK(cudaDeviceSynchronize());
        long long h;
        CK(cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost));
        printf("{\"bench\": \"dfma_latency\", \"cycles_per_dependent_dfma\": %.3f}\n", (double)h / (1000.0 * 64));
    }
    // ---- write bandwidth (2 GiB buffer)
    const size_t bytes = (size_t)2 << 30;
    double *buf;
    CK(cudaMalloc(&buf, bytes));
    for (int mode : {0, 4, 1, 2, 3}) {
        for (int cps : {8, 16}) {
            WriteArg a{buf, bytes / 8, mode, prop.multiProcessorCount * cps};
            const float ms = time_ms(launch_write, &a, 5);
            double written = (double)bytes;
            if (mode == 2) written = (double)(bytes / 8 / 120) * 32 * 8;
            if (mode == 3) written = (double)bytes / 2;
            printf("{\"bench\": \"write\", \"mode\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"GBps_written\": %.1f, \"GBps_span\": %.1f}\n", mode, cps, ms,
                   written / ms * 1e-6, (double)bytes / ms * 1e-6);
        }
    }
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (int sz : {64, 128, 256, 512, 1024, 4096}) {
        for (int cps : {2, 4}) {
            BulkArg a{(char *)buf, bytes, sz, prop.multiProcessorCount * cps};
            const float ms = time_ms(launch_bulk, &a, 5);
            printf("{\"bench\": \"bulk_store\", \"bytes_per_op\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"GBps\": %.1f, \"Mops_per_s_per_sm\": %.2f}\n", sz, cps, ms,
                   (double)bytes / ms * 1e-6, (double)bytes / sz / ms * 1e-3 / prop.multiProcessorCount);
        }
    }
    // ---- instruction cache
    const int n_of[3] = {1024, 4096, 16384};
    for (int which = 0; which < 3; ++which)
        for (int warps : {8, 16})
            for (int stag = 0; stag < 2; ++stag) {
                const int reps = (1 << 20) / n_of[which];
                // staggered: the warps of an SM are spread evenly over one pass of the body
                const int per_smsp = warps / 4;
                const int delay = stag ? (int)((long long)n_of[which] * 2 * per_smsp / warps) : 0;
                IcArg a{out, which, warps, reps, delay};
                const float ms = time_ms(launch_ic, &a, 3);
                const double inst = (double)148 * warps * (1 << 20);
                printf("{\"bench\": \"icache\", \"body_instr\": %d, \"body_KB\": %d, \"warps_per_sm\": %d, \"staggered\": %d, \"ms\": %.3f, \"dfma_per_clk_per_sm\": %.3f}\n",
                       n_of[which], n_of[which] * 16 / 1024, warps, stag, ms, inst / 148 / (ms * 1e-3 * clk_khz * 1e3));
            }
    return 0;
}
