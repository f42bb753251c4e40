/* awegpu.h — C ABI of the B200 evaluator for awebox's collocation constraint residual g(w,p) and
 * its sparse Jacobian.
 *
 * Two groups of entry points:
 *
 * (1) the batched ABI (plain pointers and sizes; what a ctypes / cgo / JNI stub binds).  It replaces
 *     the numeric evaluation of the reference's ``NLP.g_fun`` (awebox/ocp/nlp.py:119-121, 270-272)
 *     and of the ``jac_g`` that CasADi derives from it inside ``cas.nlpsol``
 *     (awebox/opti/preparation.py:366-380, 395), for B independent OCP instances per call.
 *
 * (2) the CasADi "external function" ABI (``casadi.external(name, lib)`` loads a shared library
 *     that exports ``<name>``, ``<name>_n_in``, ``<name>_sparsity_in`` ...; casadi/core/external.cpp,
 *     the same symbols ``Function.generate`` emits).  ``awegpu_g`` is g(V,P); ``jac_awegpu_g`` is the
 *     Jacobian function with CasADi 3.6 conventions (inputs V, P, out_g; outputs jac_g_V [sparse CCS],
 *     jac_g_P [structurally empty]).  They are bound to a handle with ``awegpu_casadi_bind``.
 *
 * All matrices are FP64; index arrays are int64 (``casadi_int``) on the ABI, int32 inside.
 * Return value 0 = ok, non-zero = failure (message via awegpu_last_error).  A handle owns one CUDA
 * stream and is not re-entrant.  There is NO CPU fallback: every entry point fails when no CUDA
 * device is usable.
 */
#ifndef AWEGPU_H
#define AWEGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Discretised-OCP description produced by awebox_b200/ocp/nlp.py (host side, build time).  Layout
 * follows awebox/ocp/var_struct.py:39-97 (V), awebox/ocp/discretization.py:157-179 (P) and
 * awebox/ocp/constraints.py:282-352 (rows of g).  All table pointers are host pointers; create()
 * copies them to the device. */
typedef struct awegpu_problem {
    int32_t n_v, n_p, n_g;          /* |V|, |P|, |g| */
    int64_t nnz;                    /* nnz of jac_g (w.r.t. V) */
    int32_t n_k, d;                 /* intervals, collocation order */
    int32_t n_x, n_u, n_z, n_theta, n_phi, n_theta0, n_var;
    int32_t v_phi;                  /* offset of phi in V */
    int32_t p_theta0;               /* offset of theta0 in P */
    int32_t n_coll_out, n_shoot_out;/* node-local output counts (must match the compiled model) */
    int32_t n_shoot_eq, n_ineq;
    int32_t n_const;                /* constant Jacobian entries (continuity, periodicity, affine rows) */
    int32_t n_periodic;
    int32_t row_periodic;           /* first periodicity row */
    int32_t x_last;                 /* V offset of coll_var[n_k-1, d-1].x */
    const int32_t *x_base;          /* [n_k*(d+1)]  V offset of X_r of interval k */
    const int32_t *x_next;          /* [n_k]        V offset of x_{k+1} */
    const int32_t *u_base, *xdot_base, *z_base;   /* [n_k] */
    const int32_t *cz_base;         /* [n_k*d]      V offset of coll_var[k,j].z */
    const int32_t *theta_idx;       /* [n_k*n_theta] V index of every theta entry seen by interval k */
    const int32_t *tf_idx;          /* [n_k]        V index of the t_f used by interval k */
    const int32_t *row_shoot, *row_path, *row_cont;   /* [n_k] */
    const int32_t *row_coll;        /* [n_k*d] */
    const int32_t *periodic_idx;    /* [n_periodic] state components tied by periodicity */
    const int32_t *coll_slot;       /* [n_k*d*n_coll_out]  CCS slot of every local output, -1 = none */
    const int32_t *shoot_slot;      /* [n_k*n_shoot_out] */
    const int32_t *const_slot;      /* [n_const] */
    const double  *const_val;       /* [n_const] */
    const int32_t *const_begin;     /* [n_k+1] range of constant entries written by interval k */
    const double  *coeff_coll;      /* [(d+1)*(d+1)]  C[r*(d+1)+j] = l_r'(tau_j) */
    const double  *coeff_cont;      /* [d+1]          l_r(1) */
    const int64_t *colind;          /* [n_v+1] */
    const int64_t *row;             /* [nnz] */
    /* affine rows  g[aff_row[n]] = scale * (sum_{i in [aff_begin[n], aff_begin[n+1])} aff_coef[i] * w[aff_col[i]]
     *                                      + sum_{i in [affp_begin[n], affp_begin[n+1])} affp_coef[i] * p[affp_col[i]]
     *                                      + aff_const[n]):
     * the t_f bound rows of phase_fix='single_reelout' (awebox/ocp/constraints.py:147-168) and the initial
     * conditions x_0 - p.ref[x_0] of an mpc trial (awebox/ocp/operation.py:303-326).  Rows with parameter terms
     * need the full parameter vector (not AWEGPU_P_THETA0). */
    int32_t n_aff;
    const int32_t *aff_row;         /* [n_aff] */
    const int32_t *aff_begin;       /* [n_aff+1] */
    const int32_t *aff_col;         /* [aff_begin[n_aff]]  index into V */
    const double  *aff_coef;        /* [aff_begin[n_aff]] */
    const int32_t *affp_begin;      /* [n_aff+1] */
    const int32_t *affp_col;        /* [affp_begin[n_aff]] index into P */
    const double  *affp_coef;       /* [affp_begin[n_aff]] */
    const double  *aff_const;       /* [n_aff] */
    double constraint_scale;
    /* output order of the collocation-node programs of the CUDA kernels (awebox_b200/ocp/nlp.py: kcoll_out): the
     * same structural entries as coll_slot, listed as "head" outputs (residual rows and the entries in the x_k, u_k,
     * theta and phi columns) followed by the node's own CCS range (the columns of X_{k,j+1} and z_{k,j}, all rows,
     * sibling nodes of the interval included) */
    int32_t n_kcoll_out, n_kcoll_head;
    const int32_t *kcoll_slot;      /* [n_k*d*n_kcoll_out]  CCS slot, -1 = none */
    /* objective f(w, p) (awebox/ocp/objective.py, trajectory type power_cycle; awebox_b200/ocp/objective.py);
     * obj_enabled = 0: the problem has no objective on this path (awegpu_eval_objective fails) */
    int32_t obj_enabled;
    const int32_t *obj_cat;         /* [n_var] cost category of every entry of the node column: 0 tracking, 1 xdot
                                     * regularisation, 2 u regularisation, 3 fictitious, 4 theta regularisation, -1 none */
    const double  *obj_norm;        /* [6] normalisation of the five categories and of the beta cost */
    const double  *obj_quad;        /* [d] quadrature weights of the collocation nodes */
    int32_t obj_n_tf;               /* time period T = sum_i obj_tf_coef[i] * V[obj_tf_idx[i]] */
    const int32_t *obj_tf_idx;      /* [obj_n_tf] */
    const double  *obj_tf_coef;     /* [obj_n_tf] */
    int32_t obj_p_weights, obj_p_cost;   /* offsets of p.weights [n_var] and of the cost block in P */
    int32_t obj_psi_idx;            /* V index of the homotopy parameter psi */
    const int32_t *obj_phi_cost;    /* [n_phi] index into the cost block of every homotopy parameter */
    int32_t obj_c_power, obj_c_tf, obj_c_beta;   /* indices into the cost block */
    const int32_t *obj_c_cat;       /* [5] ... of the five regularisation categories */
    /* obj_enabled = 2: the tracking cost of the periodic MPC (awebox/pmpc.py:304-358; trajectory type mpc,
     * awebox_b200/ocp/objective.py: MpcObjective).  The parameter vector of awegpu_eval_objective is then Pmpc's own
     * (pmpc.py:152-162), mpc_n_p entries: [x0 | ref (n_v) | u_ref | Q (n_x) | R (n_u) | P (n_x)].  Entry i of V carries
     * the term mpc_fac[i] * weight * (w_i - ref_i)^2 with weight = Q / R / P [mpc_idx[i]] or 1 by mpc_kind[i]
     * (0 none, 1 Q, 2 R, 3 P, 4 one). */
    int32_t mpc_n_p, mpc_p_ref, mpc_p_Q, mpc_p_R, mpc_p_P;
    const int32_t *mpc_kind;        /* [n_v] */
    const int32_t *mpc_idx;         /* [n_v] */
    const double  *mpc_fac;         /* [n_v] */
} awegpu_problem_t;

typedef struct awegpu_handle awegpu_handle_t;

/* flags of awegpu_eval_batch */
#define AWEGPU_JAC_INSTANCE_MAJOR 0   /* Jv[b*nnz + slot]   (what CasADi consumes per instance) */
#define AWEGPU_JAC_BATCH_MINOR    1   /* Jv[slot*B + b]     (for batched on-device consumers;   */
                                      /* not compiled into the libraries of very large models:  */
                                      /* the call then fails with a message)                    */
#define AWEGPU_P_FULL             0   /* P holds full parameter vectors, leading dimension n_p  */
#define AWEGPU_P_THETA0           4   /* P holds theta0 only, leading dimension n_theta0        */
#define AWEGPU_P_SHARED           8   /* one parameter vector shared by all instances           */
#define AWEGPU_SKIP_JAC          16   /* residual only                                          */
#define AWEGPU_HOST_PINNED       32   /* eval_batch_host: caller buffers are page-locked; copy   */
                                      /* directly, chunked over two streams (no staging copy)    */

const char *awegpu_model_name(void);             /* spec the node functions were generated from */
int awegpu_model_dims(int32_t *n_var, int32_t *n_param, int32_t *n_coll_out, int32_t *n_shoot_out,
                      int64_t *flops_coll, int64_t *flops_shoot);

int  awegpu_create(const awegpu_problem_t *problem, int device, awegpu_handle_t **out);
void awegpu_destroy(awegpu_handle_t *h);
const char *awegpu_last_error(awegpu_handle_t *h);   /* h may be NULL: last create() error */

int awegpu_sparsity(awegpu_handle_t *h, int64_t *colind /*[n_v+1]*/, int64_t *row /*[nnz]*/);

/* Allocate the side buffers for batches of up to B instances now (and, with AWEGPU_RESERVE_HOST, the device / pinned
 * staging buffers of awegpu_eval_batch_host; add AWEGPU_HOST_PINNED if the caller's own buffers are page-locked), so
 * that no evaluation call allocates.  create() reserves for one instance (the CasADi entry points). */
#define AWEGPU_RESERVE_HOST 64
int awegpu_reserve(awegpu_handle_t *h, int32_t B, int32_t flags);

/* Device-pointer entry: W [B,n_v], P (see flags), G [B,n_g], Jv [B*nnz] are device pointers.
 * stream: a cudaStream_t (NULL = the handle's own stream).  Asynchronous w.r.t. the host. */
int awegpu_eval_batch(awegpu_handle_t *h, int32_t B, const double *W, const double *P,
                      double *G, double *Jv, int32_t flags, void *stream);

/* Host-pointer entry: copies inputs host->device and results device->host through pinned staging
 * buffers owned by the handle (grown on demand), synchronises before returning. */
int awegpu_eval_batch_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P,
                           double *G, double *Jv, int32_t flags);

/* Device-resident products of a batch for an on-device consumer (a batched solver; the sequel to
 * awebox/sweep.py:117-204 and awebox/pmpc.py:221-302), O(n_v + n_g) values per instance instead of O(nnz):
 *   JTl   [B, n_v] = J_b^T lambda_b      lambda [B, n_g] (or one vector with AWEGPU_RED_LAMBDA_SHARED)
 *   Jv    [B, n_g] = J_b v_b             v [B, n_v]      (or one vector with AWEGPU_RED_V_SHARED)
 *   cnorm [B, n_v], rnorm [B, n_g]       max |J_b| per column / per row
 *   ginf  [B]                            max |g_b|
 * G [B, n_g] and Jvals [B, nnz] are what awegpu_eval_batch wrote (instance-major); all pointers are device
 * pointers; any output (and lambda / v with the outputs that need them) may be NULL. */
#define AWEGPU_RED_LAMBDA_SHARED 1
#define AWEGPU_RED_V_SHARED      2
int awegpu_reduce_batch(awegpu_handle_t *h, int32_t B, const double *G, const double *Jvals, const double *lambda,
                        const double *v, int32_t red_flags, double *JTl, double *Jv, double *cnorm, double *rnorm,
                        double *ginf, void *stream);
/* Host-pointer entry (pinned or pageable buffers): evaluates g and jac_g on the device, keeps the Jacobian values
 * there and returns g [B, n_g] and the requested products only.  flags: as awegpu_eval_batch (instance-major). */
int awegpu_eval_batch_reduce_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P,
                                  const double *lambda, const double *v, int32_t flags, int32_t red_flags,
                                  double *G, double *JTl, double *Jv, double *cnorm, double *rnorm, double *ginf);

/* Objective of the discretised OCP and its gradient, B instances per call (replaces the numeric evaluation of
 * NLP.f_fun and of the gradient cas.nlpsol derives from it: awebox/ocp/nlp.py:112-118, awebox/ocp/objective.py:422-520;
 * trajectory type power_cycle).  W [B, n_v]; P full parameter vectors [B, n_p] or one vector with AWEGPU_P_SHARED
 * (the objective reads p.ref, p.weights, the cost factors and theta0); F [B]; GradF [B, n_v] dense;
 * comp [B, 10] (optional): tracking, xdot regularisation, u regularisation, fictitious, theta regularisation, beta,
 * time, power, homotopy, general problem cost.  Device pointers / host pointers.
 * Trajectory type mpc (obj_enabled = 2): the tracking cost of awebox/pmpc.py:304-358; P are Pmpc's parameter vectors
 * [B, mpc_n_p] or one vector with AWEGPU_P_SHARED; comp must be NULL. */
int awegpu_eval_objective(awegpu_handle_t *h, int32_t B, const double *W, const double *P, int32_t flags,
                          double *F, double *GradF, double *comp, void *stream);
int awegpu_eval_objective_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P, int32_t flags,
                               double *F, double *GradF, double *comp);

/* Second-order hook: upper triangle, CCS, of the Hessian of the Lagrangian,
 * d^2/dw^2 (sigma f(w, p) + lam' g(w, p)), B instances per call.  Replaces the nlp_hess_l function CasADi derives inside cas.nlpsol when the
 * reference solves with the exact Hessian (awebox/opts/default.py:323 hessian_approximation = False; the Lagrangian:
 * awebox/tools/debug_operations.py:567-596; g: awebox/ocp/constraints.py:210-373).
 * awegpu_hess_setup uploads the tables of awebox_b200/ocp/hessian.py once per handle: CCS pattern of triu(H) (colind [n_v+1],
 * row [nnz]), the assembly matrix (asm_ptr [nnz+1], asm_idx / asm_coef [n_asm]: node Hessian values and coefficients of
 * every slot), V offset of x_0, first g row and rows per interval, collocation coefficients C[r, j] ((d+1) x (d+1), row-major)
 * and quadrature weights [d], entries per collocation / shooting node Hessian and whether they carry the objective part
 * (checked against the compiled programs), the time period T = sum_i t_coef[i] V[t_idx[i]] (n_t <= 2 terms).
 * W [B, n_v]; P as awegpu_eval_batch (flags AWEGPU_P_*); Lam [B, n_g] multipliers of g; Sigma [B] multiplier of f per instance
 * (nlpsol's lam_f), or NULL for the constraint part alone -- with Sigma, P must hold full parameter vectors (the objective
 * reads p.ref, p.weights and the cost factors; awebox/ocp/objective.py:422-455) and the problem must have an objective on
 * this path (trajectory type power_cycle); H [B, nnz] instance-major.
 * Device pointers (asynchronous, stream as awegpu_eval_batch) / host pointers (synchronous). */
int awegpu_hess_setup(awegpu_handle_t *h, int64_t nnz, int32_t n_hv, int64_t n_asm, const int64_t *colind,
                      const int64_t *row, const int64_t *asm_ptr, const int64_t *asm_idx, const double *asm_coef,
                      int32_t v_x0, int32_t row0, int32_t row_stride, const double *coeff, const double *quad,
                      int32_t n_ec, int32_t n_es, int32_t with_objective, int32_t n_t, const int32_t *t_idx,
                      const double *t_coef);
int awegpu_hess_has_objective(void);               /* 1: the node programs of this library carry the objective part */
int64_t awegpu_hess_nnz(awegpu_handle_t *h);       /* -1 before awegpu_hess_setup */
int awegpu_hess_sparsity(awegpu_handle_t *h, int64_t *colind /*[n_v+1]*/, int64_t *row /*[nnz]*/);
int awegpu_eval_hessian(awegpu_handle_t *h, int32_t B, const double *W, const double *P, const double *Lam,
                        const double *Sigma, int32_t flags, double *H, void *stream);
int awegpu_eval_hessian_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P, const double *Lam,
                             const double *Sigma, int32_t flags, double *H);
/* device time of the most recent awegpu_eval_hessian (ms, CUDA events on the launching stream); <0 if none */
float awegpu_last_hessian_ms(awegpu_handle_t *h);

/* ---- outputs map.  Replaces nlp.Outputs_fun(V, P) (awebox/ocp/nlp.py:335-336), the
 * horizontal concatenation of model.outputs_fun over the nodes of every interval: shooting node k, then the d
 * collocation nodes (awebox/ocp/collocation.py:366-381, awebox/ocp/discretization.py:316-334; zoh controls).
 * Out [B][n_out][n_nodes], n_nodes = n_k (d+1): Out[b] is the C-contiguous image of np.array(Outputs_fun(V_b, P_b))
 * (one label = one contiguous series over the nodes).  The reference addresses entries BY LABEL
 * ('[group,name,i]' of model.outputs); its index order depends on Python set iteration and is not fixed, so the label
 * table is part of the interface: awegpu_outputs_label(i), i < awegpu_outputs_n().
 * P: as for awegpu_eval_batch (AWEGPU_P_FULL / AWEGPU_P_THETA0 / AWEGPU_P_SHARED). */
int32_t awegpu_outputs_n(void);
const char *awegpu_outputs_label(int32_t i);
int32_t awegpu_outputs_nodes(awegpu_handle_t *h);
int awegpu_eval_outputs(awegpu_handle_t *h, int32_t B, const double *W, const double *P, int32_t flags, double *Out,
                        void *stream);
int awegpu_eval_outputs_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P, int32_t flags, double *Out);
/* device time of the most recent awegpu_eval_outputs (ms, CUDA events on the launching stream); <0 if none */
float awegpu_last_outputs_ms(awegpu_handle_t *h);

/* device time of the most recent eval (ms, CUDA events on the launching stream); <0 if none */
float awegpu_last_eval_ms(awegpu_handle_t *h);
/* device time of the dominant kernel (collocation-node programs, output stage) of the first chunk of the
 * most recent eval with Jacobian (ms, CUDA events on the launching stream); <0 if none */
float awegpu_last_kernel_ms(awegpu_handle_t *h);
/* structure of the compiled node programs: groups per stage {coll s0,s1,s2, shoot s0,s1,s2}, exported
 * intermediates {coll, shoot}, FP64 operations executed per collocation / shooting node (all stages,
 * recomputation included) */
int awegpu_kernel_info(int32_t *n_groups /*[6]*/, int32_t *n_exports /*[2]*/, int64_t *ops_coll, int64_t *ops_shoot);
/* number of kernel launches issued by this handle so far */
int64_t awegpu_launch_count(awegpu_handle_t *h);
int awegpu_synchronize(awegpu_handle_t *h);

/* ---- CasADi external-function ABI (casadi_real = double, casadi_int = long long) ---- */
/* awegpu_casadi_bind(h) makes h the handle of the Functions that are created from this library next:
 * casadi.external looks up <name>_checkout (casadi/core/external.cpp), which returns the slot of the handle bound at
 * that time, and passes it as `mem` to every evaluation, so two OCPs of one model (two handles) can live in one
 * process: bind A, create external A, bind B, create external B.  Evaluations of one handle are serialised.
 * jac_awegpu_g declares jac_g_P structurally empty (nlpsol differentiates g w.r.t. V only): g DOES depend on P, this
 * library is not a source of parametric sensitivities. */
int awegpu_casadi_bind(awegpu_handle_t *h);

int  awegpu_g_checkout(void);
void awegpu_g_release(int mem);
void awegpu_g_incref(void);
void awegpu_g_decref(void);
int  awegpu_g_alloc_mem(void);
int  awegpu_g_init_mem(int mem);
void awegpu_g_free_mem(int mem);
int  jac_awegpu_g_checkout(void);
void jac_awegpu_g_release(int mem);
void jac_awegpu_g_incref(void);
void jac_awegpu_g_decref(void);
int  jac_awegpu_g_alloc_mem(void);
int  jac_awegpu_g_init_mem(int mem);
void jac_awegpu_g_free_mem(int mem);

int awegpu_g(const double **arg, double **res, long long *iw, double *w, int mem);
long long awegpu_g_n_in(void);
long long awegpu_g_n_out(void);
const char *awegpu_g_name_in(long long i);
const char *awegpu_g_name_out(long long i);
const long long *awegpu_g_sparsity_in(long long i);
const long long *awegpu_g_sparsity_out(long long i);
int awegpu_g_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w);

int jac_awegpu_g(const double **arg, double **res, long long *iw, double *w, int mem);
long long jac_awegpu_g_n_in(void);
long long jac_awegpu_g_n_out(void);
const char *jac_awegpu_g_name_in(long long i);
const char *jac_awegpu_g_name_out(long long i);
const long long *jac_awegpu_g_sparsity_in(long long i);
const long long *jac_awegpu_g_sparsity_out(long long i);
int jac_awegpu_g_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w);

/* hess_lag for cas.nlpsol('solver', 'ipopt', nlp, {'hess_lag': casadi.external('awegpu_hess_lag', lib), ...}): inputs
 * (x, p, lam_f, lam_g), output triu of the Hessian of lam_f f + lam_g' g with the pattern of awegpu_hess_sparsity (available
 * after awegpu_hess_setup).  Problems without an objective on this path (MPC trial): constraint part, lam_f not used. */
int  awegpu_hess_lag_checkout(void);
void awegpu_hess_lag_release(int mem);
void awegpu_hess_lag_incref(void);
void awegpu_hess_lag_decref(void);
int  awegpu_hess_lag_alloc_mem(void);
int  awegpu_hess_lag_init_mem(int mem);
void awegpu_hess_lag_free_mem(int mem);
int awegpu_hess_lag(const double **arg, double **res, long long *iw, double *w, int mem);
long long awegpu_hess_lag_n_in(void);
long long awegpu_hess_lag_n_out(void);
const char *awegpu_hess_lag_name_in(long long i);
const char *awegpu_hess_lag_name_out(long long i);
const long long *awegpu_hess_lag_sparsity_in(long long i);
const long long *awegpu_hess_lag_sparsity_out(long long i);
int awegpu_hess_lag_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w);

/* objective as CasADi externals (replace NLP.f_fun, awebox/ocp/nlp.py:119-121, 142): casadi.external('awegpu_f', lib) with
 * inputs (V, P) and the scalar output f; casadi.external finds its Jacobian function jac_awegpu_f in the same library
 * (inputs V, P, out_f; outputs jac_f_V dense 1 x n_v, jac_f_P structurally empty).  Power-cycle trials: P is the full
 * parameter vector; mpc trials: Pmpc's own parameter vector (awebox/pmpc.py:152-162). */
int  awegpu_f_checkout(void);
void awegpu_f_release(int mem);
void awegpu_f_incref(void);
void awegpu_f_decref(void);
int  awegpu_f_alloc_mem(void);
int  awegpu_f_init_mem(int mem);
void awegpu_f_free_mem(int mem);
int awegpu_f(const double **arg, double **res, long long *iw, double *w, int mem);
long long awegpu_f_n_in(void);
long long awegpu_f_n_out(void);
const char *awegpu_f_name_in(long long i);
const char *awegpu_f_name_out(long long i);
const long long *awegpu_f_sparsity_in(long long i);
const long long *awegpu_f_sparsity_out(long long i);
int awegpu_f_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w);

int  jac_awegpu_f_checkout(void);
void jac_awegpu_f_release(int mem);
void jac_awegpu_f_incref(void);
void jac_awegpu_f_decref(void);
int  jac_awegpu_f_alloc_mem(void);
int  jac_awegpu_f_init_mem(int mem);
void jac_awegpu_f_free_mem(int mem);
int jac_awegpu_f(const double **arg, double **res, long long *iw, double *w, int mem);
long long jac_awegpu_f_n_in(void);
long long jac_awegpu_f_n_out(void);
const char *jac_awegpu_f_name_in(long long i);
const char *jac_awegpu_f_name_out(long long i);
const long long *jac_awegpu_f_sparsity_in(long long i);
const long long *jac_awegpu_f_sparsity_out(long long i);
int jac_awegpu_f_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w);

/* outputs map as a CasADi external (replaces NLP.Outputs_fun, awebox/ocp/nlp.py:335-336): inputs (V, P), output the dense
 * n_out x n_nodes matrix in CasADi's column-major non-zero order; rows in the order of awegpu_outputs_label. */
int  awegpu_outputs_checkout(void);
void awegpu_outputs_release(int mem);
void awegpu_outputs_incref(void);
void awegpu_outputs_decref(void);
int  awegpu_outputs_alloc_mem(void);
int  awegpu_outputs_init_mem(int mem);
void awegpu_outputs_free_mem(int mem);
int awegpu_outputs(const double **arg, double **res, long long *iw, double *w, int mem);
long long awegpu_outputs_n_in(void);
long long awegpu_outputs_n_out(void);
const char *awegpu_outputs_name_in(long long i);
const char *awegpu_outputs_name_out(long long i);
const long long *awegpu_outputs_sparsity_in(long long i);
const long long *awegpu_outputs_sparsity_out(long long i);
int awegpu_outputs_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w);

#ifdef __cplusplus
}
#endif
#endif /* AWEGPU_H */
