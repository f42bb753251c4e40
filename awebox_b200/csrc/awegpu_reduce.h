// awegpu_reduce.h — interface between the C ABI (awegpu_kernels.cu) and the batch reduction kernel (awegpu_reduce.cu)
#ifndef AWEGPU_REDUCE_H
#define AWEGPU_REDUCE_H
#include <cuda_runtime.h>

struct ReduceArgs {
    int B, n_v, n_g;
    long long nnz;
    const int *colind, *row, *col;  // CCS pattern, int32, device (col: column of every entry)
    const double *G, *J, *lambda, *v;
    long long ld_lambda, ld_v;      // 0: one vector shared by the batch
    double *JTl, *Jv, *cnorm, *rnorm, *ginf;
};

// 0 = launched; > 0 = cudaError_t
int awe_reduce_launch(cudaStream_t st, const ReduceArgs &a, int n_sm);
#endif
