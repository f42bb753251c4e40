/* awebox_oracle.c — CPU checker for the collocation g / jac_g path.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this file; the shipped package (awebox_b200/) never links or loads it.
 *
 * It is a plain-C, scalar, one-instance-at-a-time restatement of what one evaluation of the
 * reference's NLP.g_fun and of its CasADi-derived Jacobian computes:
 *   - gather of the node input column [x; xdot; u; z; theta] and [theta0; phi]
 *       awebox/tools/struct_operations.py:69-116, 363-388, 436-471
 *   - xdot from the collocation polynomial,  xdot_kj = (sum_r C[r,j] X_r) / h / t_f
 *       awebox/ocp/collocation.py:202-258
 *   - model equalities at the n_k*d collocation nodes, equality subset + inequalities at the n_k
 *     shooting nodes        awebox/ocp/constraints.py:239-268   (node arithmetic: node_*.inc)
 *   - continuity and periodicity rows   awebox/ocp/collocation.py:319-336, ocp/operation.py:245-267
 *   - row order of g        awebox/ocp/constraints.py:282-352
 *   - jac_g in CCS order (what cas.nlpsol derives, awebox/opti/preparation.py:379)
 * The node arithmetic (node_coll.inc / node_shoot.inc) is emitted from the expression graph of
 * awebox_b200 and is pinned against the reference's own Python run over a restated CasADi
 * (oracle/make_golden.py -> tests/golden/).  Parity against a live casadi==3.6.4 is UNPINNED: that
 * package cannot be installed in the build container.
 *
 * Compile:  gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC -I<gen dir> -Iinclude oracle/awebox_oracle.c
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "awegpu.h"
#include "model_defs.h"

#define RSQRT(x) (1.0 / sqrt(x))
#define DIVF(a, b) ((a) / (b))
#define SQRTF(x) sqrt(x)
#define SIGN(x) (((x) > 0.0) - ((x) < 0.0))

#define VIN(i) v[i]
#define PIN(i) p[i]
#define CCIN(r) cc[r]
#define OWNIN(r) own[r]
#define ITF itf
#define OUT(n, val) out[n] = (val)

static void node_coll(const double *v, const double *p, const double *cc, const double *own, double itf,
                      double *out)
{
#include "node_coll.inc"
}

static void node_shoot(const double *v, const double *p, double *out)
{
#include "node_shoot.inc"
}

const char *awebox_oracle_model_name(void) { return AWE_MODEL_NAME; }

/* one instance: w [n_v], theta0 [n_theta0] -> g [n_g], jv [nnz] (jv may be NULL) */
static void eval_one(const awegpu_problem_t *pr, const double *w, const double *theta0, const double *pfull, double *g,
                     double *jv)
{
    const int d = pr->d, n_x = pr->n_x;
    double v[AWE_N_VAR], p[AWE_N_PARAM], out_c[AWE_N_COLL_OUT], out_s[AWE_N_SHOOT_OUT];
    double cc[AWE_D + 1], own[AWE_D + 1];
    int k, j, r, m, n, t;

    for (n = 0; n < pr->n_theta0; ++n) p[n] = theta0[n];
    for (n = 0; n < pr->n_phi; ++n) p[pr->n_theta0 + n] = w[pr->v_phi + n];

    for (k = 0; k < pr->n_k; ++k) {
        const int32_t *xb = pr->x_base + k * (d + 1);
        const double tf = w[pr->tf_idx[k]];
        /* ---- shooting node: x_k, lifted xdot_k, u_k, z_k, theta */
        for (m = 0; m < n_x; ++m) v[m] = w[xb[0] + m];
        for (m = 0; m < n_x; ++m) v[n_x + m] = w[pr->xdot_base[k] + m];
        for (m = 0; m < pr->n_u; ++m) v[2 * n_x + m] = w[pr->u_base[k] + m];
        for (m = 0; m < pr->n_z; ++m) v[2 * n_x + pr->n_u + m] = w[pr->z_base[k] + m];
        for (t = 0; t < pr->n_theta; ++t) v[2 * n_x + pr->n_u + pr->n_z + t] = w[pr->theta_idx[k * pr->n_theta + t]];
        node_shoot(v, p, out_s);
        for (n = 0; n < pr->n_shoot_eq; ++n) g[pr->row_shoot[k] + n] = out_s[n];
        for (n = 0; n < pr->n_ineq; ++n) g[pr->row_path[k] + n] = out_s[pr->n_shoot_eq + n];
        if (jv) {
            const int32_t *slot = pr->shoot_slot + (size_t)k * pr->n_shoot_out;
            for (n = pr->n_shoot_eq + pr->n_ineq; n < pr->n_shoot_out; ++n)
                if (slot[n] >= 0) jv[slot[n]] = out_s[n];
        }
        /* ---- collocation nodes */
        for (j = 0; j < d; ++j) {
            const double h = 1.0 / pr->n_k;
            for (m = 0; m < n_x; ++m) {
                double xp = 0.0;
                for (r = 0; r <= d; ++r) xp += pr->coeff_coll[r * (d + 1) + (j + 1)] * w[xb[r] + m];
                v[m] = w[xb[j + 1] + m];
                v[n_x + m] = xp / h / tf;
            }
            for (m = 0; m < pr->n_z; ++m) v[2 * n_x + pr->n_u + m] = w[pr->cz_base[k * d + j] + m];
            for (r = 0; r <= d; ++r) {
                cc[r] = pr->coeff_coll[r * (d + 1) + (j + 1)] / h / tf;
                own[r] = (r == j + 1) ? 1.0 : 0.0;
            }
            node_coll(v, p, cc, own, 1.0 / tf, out_c);
            for (n = 0; n < AWE_N_EQ; ++n) g[pr->row_coll[k * d + j] + n] = out_c[n];
            if (jv) {
                const int32_t *slot = pr->coll_slot + ((size_t)k * d + j) * pr->n_coll_out;
                for (n = AWE_N_EQ; n < pr->n_coll_out; ++n)
                    if (slot[n] >= 0) jv[slot[n]] = out_c[n];
            }
        }
        /* ---- continuity: x_{k+1} - sum_r D[r] X_r */
        for (m = 0; m < n_x; ++m) {
            double xf = 0.0;
            for (r = 0; r <= d; ++r) xf += pr->coeff_cont[r] * w[xb[r] + m];
            g[pr->row_cont[k] + m] = pr->constraint_scale * (w[pr->x_next[k] + m] - xf);
        }
    }
    for (n = 0; n < pr->n_periodic; ++n)
        g[pr->row_periodic + n] = pr->constraint_scale *
                                  (w[pr->x_base[0] + pr->periodic_idx[n]] - w[pr->x_last + pr->periodic_idx[n]]);
    /* ---- affine rows (t_f bounds of single_reelout, awebox/ocp/constraints.py:147-168; initial conditions of an
     *      mpc trial, awebox/ocp/operation.py:303-326) */
    for (n = 0; n < pr->n_aff; ++n) {
        double a = 0.0;
        for (r = pr->aff_begin[n]; r < pr->aff_begin[n + 1]; ++r) a += pr->aff_coef[r] * w[pr->aff_col[r]];
        for (r = pr->affp_begin[n]; r < pr->affp_begin[n + 1]; ++r) a += pr->affp_coef[r] * pfull[pr->affp_col[r]];
        g[pr->aff_row[n]] = pr->constraint_scale * (a + pr->aff_const[n]);
    }
    if (jv)
        for (n = 0; n < pr->n_const; ++n) jv[pr->const_slot[n]] = pr->const_val[n];
}

/* B instances; W [B,n_v]; P per flags (AWEGPU_P_*); G [B,n_g]; Jv [B,nnz] instance-major or NULL.
 * n_threads <= 0: all OpenMP threads. */
int awebox_oracle_eval(const awegpu_problem_t *pr, int B, const double *W, const double *P, double *G,
                       double *Jv, int flags, int n_threads)
{
    long b;
    if (pr->n_coll_out != AWE_N_COLL_OUT || pr->n_shoot_out != AWE_N_SHOOT_OUT || pr->n_var != AWE_N_VAR ||
        pr->n_theta0 != AWE_N_THETA0 || pr->d != AWE_D)
        return 1;
    if ((flags & AWEGPU_P_THETA0) && pr->n_aff > 0 && pr->affp_begin[pr->n_aff] > 0) return 2;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
#endif
    for (b = 0; b < B; ++b) {
        const double *p, *pfull = NULL;
        if (flags & AWEGPU_P_SHARED)
            p = (flags & AWEGPU_P_THETA0) ? P : P + pr->p_theta0;
        else
            p = (flags & AWEGPU_P_THETA0) ? P + (size_t)b * pr->n_theta0 : P + (size_t)b * pr->n_p + pr->p_theta0;
        if (!(flags & AWEGPU_P_THETA0)) pfull = (flags & AWEGPU_P_SHARED) ? P : P + (size_t)b * pr->n_p;
        eval_one(pr, W + (size_t)b * pr->n_v, p, pfull, G + (size_t)b * pr->n_g,
                 (Jv && !(flags & AWEGPU_SKIP_JAC)) ? Jv + (size_t)b * pr->nnz : NULL);
    }
    return 0;
}
