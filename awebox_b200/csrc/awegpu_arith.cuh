// awegpu_arith.cuh — FP64 reciprocal / division / root sequences shared by the generated node programs
// (awegpu_kernels.cu, awegpu_hess.cu).  See the comment above the include in awegpu_kernels.cu for the contract.
#ifndef AWEGPU_ARITH_CUH
#define AWEGPU_ARITH_CUH
#define SIGN(x) ((double)(((x) > 0.0) - ((x) < 0.0)))
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
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fmin(y, 1e300);                         // rsqrt(0) = inf: keep 0 * y = 0
    double e = fma(-x * y, y, 1.0);
    y = fma(y * 0.5, e, y);
    e = fma(-x * y, y, 1.0);
    y = fma(y * 0.5, e, y);
    const double s = x * y;
    return fma(fma(-s, s, x), 0.5 * y, s);
}
#define RSQRT(x) rsqrt_nr(x)
#define SQRTF(x) sqrt_nr(x)
#define DIVF(a, b) div_nr(a, b)
#endif
