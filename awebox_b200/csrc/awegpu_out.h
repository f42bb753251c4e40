// awegpu_out.h — interface between the C ABI (awegpu_kernels.cu) and the kernel of the outputs map (awegpu_out.cu)
#ifndef AWEGPU_OUT_H
#define AWEGPU_OUT_H
#include <cuda_runtime.h>

struct OutArgs {
    int B, n_v, n_k;
    const double *W, *P;           // [B, n_v], parameter vectors
    long long ld_p;                // leading dimension of P (0: shared)
    int p_off;                     // theta0 of instance b: P + b * ld_p + p_off
    int v_x0, v_phi;               // V offsets of x_0 and phi
    const int *theta_idx;          // [n_k][n_theta] V index of every theta entry of interval k
    const int *tf_idx;             // [n_k]
    const double *coeff;           // [(d+1)*(d+1)] C[r*(d+1)+j] (device)
    double *Out;                   // [B][n_out][n_k*(d+1)]
};

// 0 = launched, > 0 = cudaError_t
int awe_out_launch(cudaStream_t st, const OutArgs &a);
int awe_out_n(void);
const char *awe_out_label(int i);
#endif
