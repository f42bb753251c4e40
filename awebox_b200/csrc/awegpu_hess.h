// awegpu_hess.h — interface between the C ABI (awegpu_kernels.cu) and the kernels of the second-order hook
// (awegpu_hess.cu): upper triangle of d^2/dw^2 (lam' g(w, p)) for a batch of instances.
#ifndef AWEGPU_HESS_H
#define AWEGPU_HESS_H
#include <cuda_runtime.h>

struct HessArgs {
    int B, n_v, n_g, n_k, n_hv;
    long long nnz;                 // entries of triu(H)
    const double *W, *P, *Lam;     // [B, n_v], parameter vectors, [B, n_g]
    long long ld_p;                // leading dimension of P (0: shared)
    int p_off;                     // theta0 of instance b: P + b * ld_p + p_off
    int v_x0, v_phi;               // V offsets of x_0 and phi
    int row0, row_stride;          // first g row of interval 0 (shooting rows), rows per interval
    const int *theta_idx;          // [n_k][n_theta] V index of every theta entry of interval k
    const int *asm_ptr, *asm_idx;  // [nnz + 1], [n_asm]: node values of every CCS slot (index into the per-instance vector)
    const double *asm_coef;        // [n_asm]
    // objective part (libraries generated with AWE_H_OBJ): multiplier lam_f per instance (NULL: constraint part only),
    // P must then hold full parameter vectors (p.ref | p.weights | cost | theta0); zeros: n_p zeros read instead otherwise
    const double *Sigma, *zeros;
    int n_var, n_t, t_idx[2];      // time period T = sum_i t_coef[i] * V[t_idx[i]]
    double t_coef[2];
    double *scr;                   // [tiles of this launch][n_hv][32] node Hessian values, lane-contiguous
    double *H;                     // [B, nnz]
};

// evaluates tiles [tile0, tile0 + n_tiles) of the batch; 0 = launched, > 0 = cudaError_t
int awe_hess_launch(cudaStream_t st, const HessArgs &a, int tile0, int n_tiles);
// bytes of a.scr per tile
size_t awe_hess_scratch_per_tile(int n_hv);
// collocation coefficients C[r, j], (d+1) x (d+1) row-major, into constant memory (once per process and device)
int awe_hess_set_coeff(const double *C, const double *quad);
// 1 if the node programs of this library carry the objective part; entries per collocation / shooting node Hessian
int awe_hess_has_objective(void);
void awe_hess_node_dims(int *n_ec, int *n_es);
#endif
