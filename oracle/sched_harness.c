/* sched_harness.c — CPU harness for the SCHEDULED node programs (k_*.inc).  TEST INFRASTRUCTURE.
 *
 * The CUDA kernels include k_coll.inc / k_shoot.inc (register-allocated, explicit scratch traffic,
 * staged outputs; awebox_b200/schedule.py).  This file compiles the very same text for the host with
 * array-backed macros, so that the allocator and the emission can be checked against the plain
 * emission (node_coll.inc / node_shoot.inc, used by awebox_oracle.c) without a GPU.
 *
 *   gcc -O1 -ffp-contract=off -shared -fPIC -I<gen dir> oracle/sched_harness.c
 */
#include <math.h>
#include <string.h>
#include "model_defs.h"

#define RSQRT(x) (1.0 / sqrt(x))
#define DIVF(a, b) ((a) / (b))
#define SQRTF(x) sqrt(x)
#define SIGN(x) (((x) > 0.0) - ((x) < 0.0))
#define N_D AWE_D
static const double awe_kconst[AWE_N_KCONST] = AWE_KCONST_INIT;
#define KC(i) awe_kconst[i]
#define N_X AWE_N_X

#define LD_XO(reg, m) reg = xo[m]
#define LD_XD(reg, m) reg = xd[m]
#define LD_U(reg, c) reg = u[c]
#define LD_Z(reg, c) reg = z[c]
#define LD_P(reg, i) reg = p[i]
#define LD_CC(reg, r) reg = cc[r]
#define LD_ITF(reg) reg = itf
#define LD_TH(reg, t) reg = th[t]
#define LD_OWN(reg, r) reg = (own_r == (r)) ? 1.0 : 0.0
/* inputs of the sibling nodes of the interval and the coefficients of the node's own column block
 * (awebox_b200/ocp/nlp.py: kcoll_own) */
#define LD_XS(reg, dl, m) reg = xs[((dl) - 1) * AWE_N_X + (m)]
#define LD_XDS(reg, dl, m) reg = xds[((dl) - 1) * AWE_N_X + (m)]
#define LD_ZS(reg, dl, c) reg = zs[((dl) - 1) * AWE_N_Z + (c)]
#define LD_CCO(reg) reg = cco
#define LD_CCS(reg, dl) reg = ccs[dl]
#define LEAF_FENCE()
#define LD_EX(reg, i) reg = ex_[i]
/* asynchronous input prefetch ring (awebox_b200/schedule.py: prefetch_leaves): copies complete in issue
 * order and only when a wait forces them to, so that a wrong wait count reads a stale slot */
#define PFH_XO(m)
#define PFH_XD(m)
#define PFH_U(c)
#define PFH_Z(c)
#define PFH_EX(i)
#define PF_(slot, val) do { pf_slot[pf_n % 4096] = (slot); pf_val[pf_n % 4096] = (val); ++pf_n; } while (0)
#define PF_XO(slot, m) PF_(slot, xo[m])
#define PF_XD(slot, m) PF_(slot, xd[m])
#define PF_U(slot, c) PF_(slot, u[c])
#define PF_Z(slot, c) PF_(slot, z[c])
#define PF_P(slot, i) PF_(slot, p[i])
#define PF_TH(slot, t) PF_(slot, th[t])
#define PF_EX(slot, i) PF_(slot, ex_[i])
#define LDR(reg, slot, n) do { while (pf_done < pf_n - (n)) { ring[pf_slot[pf_done % 4096]] = pf_val[pf_done % 4096]; ++pf_done; } \
                               reg = ring[slot]; } while (0)
#define EXOUT(i, val) ex_[i] = (val)
#define SPILL_ST(slot, reg) hot[slot] = reg
#define SPILL_LD(reg, slot) reg = hot[slot]
#define OVF_ST(slot, reg) cold[slot] = reg
#define OVF_LD(reg, slot) reg = cold[slot]
#define SYNC_HINT()
#define OUT(n, val) stg_[n] = (val)
#define FLUSH_PREP(q, first)
#define GROUP_INIT(q0, nq, first)
#define FLUSH(q, first, cnt, qn, fn) do { int q_; for (q_ = 0; q_ < (cnt); ++q_) out[(first) + q_] = stg_[q_]; } while (0)
/* own CCS range: results are returned in kernel output order (ocp.kcoll_out) */
#define OWN_INIT(oq0, noq, ofirst, ocnt) own_first_ = (ofirst)
#define OWN_BEGIN(q)
#define OUTO(idx, val) out[own_first_ + (idx)] = (val)
#define FLUSH_OWN(q)

/* xo [n_x] state block of the node, xd [n_x] its time derivative, u, z, th, p [n_param], cc [d+1], itf,
 * j1 = own block index in 1..d (collocation) */
#define RUN_GROUPS(FILE_, NGROUPS_)              \
    for (group = 0; group < (NGROUPS_); ++group) \
        switch (group) {                         \
            FILE_                                \
        }

void sched_coll(const double *xo, const double *xd, const double *u, const double *z, const double *th,
                const double *p, const double *cc, double itf, int own_r, double *out, int full,
                const double *xs, const double *xds, const double *zs, double cco, const double *ccs)
{
    double hot[512], cold[512], stg_[64], ex_[1024], ring[256], pf_val[4096];
    int pf_slot[4096], pf_n = 0, pf_done = 0;
    int group, own_first_ = 0;
    (void)cold, (void)own_first_;
    if (full) {
        for (group = 0; group < AWE_K_COLL_S2_GROUPS; ++group) switch (group) {
#include "k_coll_s2.inc"
            }
        for (group = 0; group < AWE_K_COLL_S1_GROUPS; ++group) switch (group) {
#include "k_coll_s1.inc"
            }
        for (group = 0; group < AWE_K_COLL_S0_GROUPS; ++group) switch (group) {
#include "k_coll_s0.inc"
            }
    } else {
        for (group = 0; group < AWE_K_COLL_RES_S2_GROUPS; ++group) switch (group) {
#include "k_coll_res_s2.inc"
            }
        for (group = 0; group < AWE_K_COLL_RES_S1_GROUPS; ++group) switch (group) {
#include "k_coll_res_s1.inc"
            }
        for (group = 0; group < AWE_K_COLL_RES_S0_GROUPS; ++group) switch (group) {
#include "k_coll_res_s0.inc"
            }
    }
}

void sched_shoot(const double *xo, const double *xd, const double *u, const double *z, const double *th,
                 const double *p, double *out, int full)
{
    double hot[512], cold[512], stg_[64], ex_[1024], ring[256], pf_val[4096];
    int pf_slot[4096], pf_n = 0, pf_done = 0;
    const double *cc = 0, *xs = 0, *xds = 0, *zs = 0, *ccs = 0;
    const double itf = 0.0, cco = 0.0;
    const int own_r = -1;
    int group, own_first_ = 0;
    (void)cc, (void)itf, (void)own_r, (void)cold, (void)xs, (void)xds, (void)zs, (void)ccs, (void)cco, (void)own_first_;
    if (full) {
        for (group = 0; group < AWE_K_SHOOT_S2_GROUPS; ++group) switch (group) {
#include "k_shoot_s2.inc"
            }
        for (group = 0; group < AWE_K_SHOOT_S1_GROUPS; ++group) switch (group) {
#include "k_shoot_s1.inc"
            }
        for (group = 0; group < AWE_K_SHOOT_S0_GROUPS; ++group) switch (group) {
#include "k_shoot_s0.inc"
            }
    } else {
        for (group = 0; group < AWE_K_SHOOT_RES_S2_GROUPS; ++group) switch (group) {
#include "k_shoot_res_s2.inc"
            }
        for (group = 0; group < AWE_K_SHOOT_RES_S1_GROUPS; ++group) switch (group) {
#include "k_shoot_res_s1.inc"
            }
        for (group = 0; group < AWE_K_SHOOT_RES_S0_GROUPS; ++group) switch (group) {
#include "k_shoot_res_s0.inc"
            }
    }
}

/* plain emission, same calling convention as awebox_oracle.c */
#undef OUT
#define VIN(i) v[i]
#define PIN(i) p[i]
#define CCIN(r) cc[r]
#define OWNIN(r) own[r]
#define ITF itf
#define OUT(n, val) out[n] = (val)
void plain_coll(const double *v, const double *p, const double *cc, const double *own, double itf, double *out)
{
#include "node_coll.inc"
}
void plain_shoot(const double *v, const double *p, double *out)
{
#include "node_shoot.inc"
}
int harness_dims(int *n_coll_out, int *n_shoot_out)
{
    *n_coll_out = AWE_N_COLL_OUT;
    *n_shoot_out = AWE_N_SHOOT_OUT;
    return 0;
}
