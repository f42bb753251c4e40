// awegpu_kernels.cu — sm_100a kernels and C ABI (include/awegpu.h) of the collocation g / jac_g
// evaluator.  Compiled once per model spec: the node arithmetic comes from the generated, grouped,
// scheduled and register-allocated k_*.inc files (awebox_b200/codegen.py, awebox_b200/schedule.py);
// everything else here is hand-written.
//
// Pipeline per call, all on one stream:
//   1. awegpu_gather_kernel   grid (tiles, n_k): stages the interval slab of 32 instances in shared
//        memory and writes the node-input records lane-contiguously (record entry i of the 32
//        instances of a tile = one 256-byte line), contracts the collocation polynomial
//        xdot_kj = (sum_r C[r,j] X_r) / (h t_f), writes the continuity / periodicity / affine rows of g and
//        the constant Jacobian entries the cooperative flush does not carry.
//   2. awegpu_node_kernel<coll | shoot, stage 2, 1>   pre-stages: exported intermediates of every node.
//   3. awegpu_node_kernel<shoot, stage 0>, awegpu_node_kernel<coll, stage 0>   the node outputs.
//      grid (ceil(items / 4), groups): warp = one node of a tile, lane = instance, blockIdx.y = program group
//      (CTAs are re-ordered so that the groups of one tile run close in time).
// A program group is a straight-line FP64 program of <= ~1400 operations that produces a consecutive
// range of the node's outputs and recomputes what it shares with the other groups.  That keeps every
// program near the 32 KB instruction cache of the SM (the monolithic ~9000-operation program was
// instruction-fetch bound) and its working set at 48 FP64 registers + a few dozen shared-memory spill slots:
// 168 registers, ~70 KB of shared memory per 4-warp CTA, three CTAs per SM.  Results are staged AWE_CHUNK
// outputs x 32 instances per warp in shared memory; the four warps of a collocation CTA are the four nodes of
// one interval and flush their chunks together, destination-sorted (flush_coop), so that neighbours in the CCS
// value array leave the SM in one store instruction.  Measured cost structure and the experiments that did not
// pay: DESIGN.md section 10.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <algorithm>
#include <utility>

#include "awegpu.h"
#include "awegpu_reduce.h"
#include "awegpu_hess.h"
#include "awegpu_out.h"
#include "model_defs.h"

// Translation units.  The generated programs make this file expensive to compile (minutes of single-threaded
// front-end time for a large model), so the build compiles it several times side by side, each time with
// -DAWE_TU=<n>, and links the objects: unit n instantiates one family of node kernels together with the functions
// that launch them, unit 0 holds the gather kernel, the residual-only kernels and the host code / C ABI.
// Without -DAWE_TU everything is in one unit.
#ifndef AWE_TU
#define AWE_TU -1
#endif
#define TU_OWNS(n) (AWE_TU == -1 || AWE_TU == (n))
#define AWE_N_TU 5      /* 0 host, 1 coll full stage 0, 2 coll full stages 1-2, 3 shoot full stage 0, 4 shoot full stages 1-2 */

#define TILE 32
#ifndef AWE_NODE_CTAS
#define AWE_NODE_CTAS 3          /* resident CTAs (of 4 warps) per SM the node kernels are compiled for */
#endif
#define N_X AWE_N_X
#define N_U AWE_N_U
#define N_Z AWE_N_Z
#define N_TH AWE_N_THETA
#define N_PHI AWE_N_PHI
#define N_D AWE_D
#define NODE_WARPS 4
#define NODE_THREADS (32 * NODE_WARPS)
#define GATHER_THREADS 128

// offsets inside the interval slab of V (awebox/ocp/var_struct.py:49-93)
#define OFF_X0 0
#define OFF_U (N_X)
#define OFF_XDOT (N_X + N_U)
#define OFF_Z (2 * N_X + N_U)
#define OFF_CX(j) (2 * N_X + N_U + N_Z + (j) * (N_X + N_Z))
#define OFF_CZ(j) (OFF_CX(j) + N_X)
#define OFF_XNEXT (2 * N_X + N_U + N_Z + N_D * (N_X + N_Z))
#define OFF_XR(r) ((r) == 0 ? OFF_X0 : OFF_CX((r) - 1))
#define SLAB_W OFF_XNEXT                    /* slab entries that go into the record */
#define SLAB_ALL (OFF_XNEXT + N_X)          /* + x_{k+1} (continuity rows) */
#define SLAB_STRIDE (SLAB_ALL | 1)
// node-input record of one (tile, interval): [slab (SLAB_W) | theta (N_TH) | phi (N_PHI)] x 32 lanes
#define REC_TH SLAB_W
#define REC_PHI (SLAB_W + N_TH)
#define REC (SLAB_W + N_TH + N_PHI)
// output staging: AWE_CHUNK outputs x 32 instances per warp, row stride odd (conflict-free both ways)
#define STG_STRIDE 33
#define STG_DOUBLES (AWE_CHUNK * STG_STRIDE)

struct DevProblem {
    int n_v, n_p, n_g, n_k, n_theta0, n_phi, v_phi, p_theta0;
    long long nnz;
    int n_coll_out, n_shoot_out, n_shoot_eq, n_ineq, n_periodic, row_periodic, x_last, x_first;
    const int *x_base, *theta_idx, *tf_idx, *row_cont, *periodic_idx;
    const int *coll_dest, *shoot_dest, *const_slot, *const_begin;
    // interval-cooperative flush of the collocation kernel: for interval k and output chunk q, the
    // N_D * AWE_CHUNK staged values of the interval's nodes sorted by destination
    const int2 *coop_tab;     // [n_k][AWE_K_COLL_NCHUNKS][N_D * AWE_CHUNK]  {destination code (-1: none, at the end),
                              //                                              staged value index  j * AWE_CHUNK + o}
    const double *const_val;
    // the same for calls whose collocation kernel flushes cooperatively: only the entries the flush tables do not carry
    const int *rest_slot, *rest_begin;
    const double *rest_val;
    const double *coop_const;      // [COOP_NCONST] distinct values of the constant entries flushed cooperatively
                                   // (per handle: several handles may share this library)
    // own CCS range of every collocation node (k, j): the columns of X_{k,j+1} and z_{k,j}, written by the thread
    // that evaluates the node, chunk by chunk (whole columns; awebox_b200/ocp/nlp.py: kcoll_own)
    const int4 *own_desc;          // [n_k * d][AWE_K_COLL_NOWNQ] {first CCS slot, entries, first constant, constants}
    const unsigned char *own_pos;  // [n_k * d][AWE_N_KCOLL_OWN]  position of every own output inside its chunk
                                   //                             (OWN_DUMP: not a structural entry)
    const uchar2 *own_cst;         // constant entries inside the own ranges {position in the chunk, value id}
    const double *own_cval;        // [OWN_NCVAL] their distinct values
    int n_aff;
    const int *aff_row, *aff_begin, *aff_col, *affp_begin, *affp_col;
    const double *aff_coef, *affp_coef, *aff_const;
    double cscale;
};

// side buffers of one chunk (device memory owned by the handle)
struct Scratch {
    double *rec;      // [tiles][n_k][REC][32]
    double *xdc;      // [tiles][n_k][d][N_X][32]
    double *pt;       // [tiles][n_theta0][32]
    double *exc;      // [tiles][n_k][d][AWE_K_COLL_EXPORTS][32]   values exported by the pre-stage programs
    double *exs;      // [tiles][n_k][AWE_K_SHOOT_EXPORTS][32]
};

// literal constants of the generated programs: read as constant-bank operands of the FP64 instructions
__constant__ double c_kconst[AWE_N_KCONST] = AWE_KCONST_INIT;
#define KC(i) c_kconst[i]
#define COOP_NCONST 8
__constant__ double c_coeff_coll[(N_D + 1) * (N_D + 1)];
__constant__ double c_coeff_cont[N_D + 1];

// Reciprocal / reciprocal square root from the MUFU seed (>= 20 bits) and two Newton steps, without the
// special-case paths of the CUDA library versions.  Division and square root get one extra residual
// correction, which leaves them within 1 ulp of the correctly rounded result for finite, normal arguments
// (norms, squared norms + eps, densities, times).  Degenerate arguments: x / 0, x / inf, rsqrt(0), sqrt(inf) and
// NaN arguments give NaN where IEEE arithmetic gives +-inf, 0 or NaN -- the same set of non-finite results except
// for x / inf -- and sqrt(0) = 0 exactly (the seed is clamped), so that a zero norm stays a zero norm; denormal
// arguments are read as zero.  Testing every division and root for these cases (a compare, a branch and an
// out-of-line IEEE sequence per operation) was measured: 7 % (AP2) to 15 % (3-DOF) of the step, and rejected;
// tests/test_gpu_parity.py::test_non_finite_and_degenerate_inputs_match_the_checker pins the behaviour.
#include "awegpu_arith.cuh"

// ---- statement macros of the generated code -------------------------------------------------------
// asm volatile: the compiler must neither merge the generator's reloads into one long-lived register
// nor hoist them.  Record entries of the 32 lanes are 256 bytes apart: base register + immediate.
#if defined(AWE_EXP_NO_LOADS)
/* ablation: inputs are not loaded (results are wrong): what the programs cost without memory latency */
#define LDG_(reg, ptr) asm volatile("mov.f64 %0, 0d3FF199999999999A;" : "=d"(reg))
#elif defined(AWE_VOLATILE_LD)
#define LDG_(reg, ptr) asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(reg) : "l"(ptr))
#else
#define LDG_(reg, ptr) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(reg) : "l"(ptr))
#endif
#ifdef AWE_VOLATILE_LD
#define LDGI_(reg, ptr) asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(reg) : "l"(ptr))
#else
#define LDGI_(reg, ptr) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(reg) : "l"(ptr))
#endif
#define LDS_(reg, addr) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(reg) : "r"(addr))
#define STS_(addr, reg) asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(reg))
// ptxas sinks loads towards their first use to save registers, which would undo the generator's
// look-ahead; it does not move memory operations across a warp barrier
#define LEAF_FENCE() asm volatile("bar.warp.sync 0xffffffff;" ::: "memory")
#define LD_XO(reg, m) LDG_(reg, w_own + (m) * 32)
#define LD_XD(reg, m) LDG_(reg, w_xd + (m) * 32)
#define LD_U(reg, c) LDG_(reg, w_rec + (OFF_U + (c)) * 32)
#define LD_Z(reg, c) LDG_(reg, w_z + (c) * 32)
#define LD_TH(reg, t) LDG_(reg, w_rec + (REC_TH + (t)) * 32)
#define LD_P(reg, i) LDG_(reg, ((i) < AWE_N_THETA0 ? w_pt + (i) * 32 : w_rec + (REC_PHI + (i) - AWE_N_THETA0) * 32))
// exports are read-only data for the per-stage kernels, but produced by the same thread inside the fused kernel
#define LD_EX(reg, i)                                                                         \
    do {                                                                                      \
        if (EX_RW)                                                                            \
            asm volatile("ld.global.f64 %0, [%1];" : "=d"(reg) : "l"(w_ex + (i) * 32) : "memory"); \
        else                                                                                  \
            LDG_(reg, w_ex + (i) * 32);                                                       \
    } while (0)
#define EXOUT(i, val) asm volatile("st.global.f64 [%0], %1;" ::"l"(w_ex + (i) * 32), "d"(val) : "memory")
#define LD_CC(reg, r) reg = c_coeff_coll[(r) * (N_D + 1) + j1] * itf_nk
// inputs of the sibling node (j + dl) mod d of the same interval and the coefficients C[j+1, .] / (h t_f) of the
// node's own column block (awebox_b200/ocp/nlp.py: kcoll_own)
#define LD_XS(reg, dl, m) LDG_(reg, w_rec + sib_off[(dl)-1] + (m) * 32)
#define LD_ZS(reg, dl, c) LDG_(reg, w_rec + sib_off[(dl)-1] + (N_X + (c)) * 32)
#define LD_XDS(reg, dl, m) LDG_(reg, w_xd + sib_xd[(dl)-1] + (m) * 32)
#define LD_CCO(reg) reg = c_coeff_coll[j1 * (N_D + 1) + j1] * itf_nk
#define LD_CCS(reg, dl) reg = c_coeff_coll[j1 * (N_D + 1) + sib_j1[(dl)-1]] * itf_nk
#define LD_ITF(reg) reg = itf
#define LD_OWN(reg, r) reg = (j1 == (r)) ? 1.0 : 0.0
#define SPILL_ST(slot, reg) STS_(scr_addr + (slot) * (8 * NODE_THREADS), reg)
#define SPILL_LD(reg, slot) LDS_(reg, scr_addr + (slot) * (8 * NODE_THREADS))
#define OVF_ST(slot, reg) SPILL_ST(slot, reg)
#define OVF_LD(reg, slot) SPILL_LD(reg, slot)
// asynchronous input prefetch (awebox_b200/schedule.py: prefetch_leaves): PF_* copies the lane's 8 bytes of
// an input from global memory into one of the thread's scratch slots without passing through a register
// (LDGSTS), one commit group per copy; LDR waits until at most n younger copies are pending (copies
// complete in issue order) and reads the slot
#define CPA_(slot, ptr)                                                                      \
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n\tcp.async.commit_group;" ::    \
                 "r"(scr_addr + (slot) * (8 * NODE_THREADS)), "l"(ptr) : "memory")
#define LDR(reg, slot, n)                                                        \
    asm volatile("cp.async.wait_group %2;\n\tld.shared.f64 %0, [%1];"            \
                 : "=d"(reg) : "r"(scr_addr + (slot) * (8 * NODE_THREADS)), "n"(n) : "memory")
// L2 prefetch hints (awebox_b200/schedule.py: prefetch_hints)
#define PFH_(ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr))
#define PFH_XO(m) PFH_(w_own + (m) * 32)
#define PFH_XD(m) PFH_(w_xd + (m) * 32)
#define PFH_U(c) PFH_(w_rec + (OFF_U + (c)) * 32)
#define PFH_Z(c) PFH_(w_z + (c) * 32)
#define PFH_EX(i) PFH_(w_ex + (i) * 32)
#define PF_XO(slot, m) CPA_(slot, w_own + (m) * 32)
#define PF_XD(slot, m) CPA_(slot, w_xd + (m) * 32)
#define PF_U(slot, c) CPA_(slot, w_rec + (OFF_U + (c)) * 32)
#define PF_Z(slot, c) CPA_(slot, w_z + (c) * 32)
#define PF_TH(slot, t) CPA_(slot, w_rec + (REC_TH + (t)) * 32)
#define PF_P(slot, i) CPA_(slot, ((i) < AWE_N_THETA0 ? w_pt + (i) * 32 : w_rec + (REC_PHI + (i) - AWE_N_THETA0) * 32))
#define PF_EX(slot, i) CPA_(slot, w_ex + (i) * 32)

// staged output: OUT parks value number n of the current chunk for this lane's instance, FLUSH writes
// chunk q (outputs first .. first+cnt-1 of the node) for the 32 instances of the tile and prefetches
// the destination table of the next chunk
// table entries per chunk: the N_D * AWE_CHUNK staged values + room for constant Jacobian entries (continuity
// rows) that lie between them in the CCS value array; COOP_NCONST rows of replicated constants follow the
// staging rows of the quad
// output stores: 0 = default write-back policy, 1 = streaming (evict-first; measured 3 % faster per step than 0 and 2), 2 = write-through
#ifndef AWE_ST_POLICY
#define AWE_ST_POLICY 1
#endif
#if AWE_ST_POLICY == 1
#define ST_OUT(ptr, val) __stcs(ptr, val)
#elif AWE_ST_POLICY == 2
#define ST_OUT(ptr, val) __stwt(ptr, val)
#else
#define ST_OUT(ptr, val) (*(ptr) = (val))
#endif
#define COOP_STAGED (N_D * AWE_CHUNK)
#define COOP_PER (COOP_STAGED + 32)
#define COOP_E ((COOP_PER + 31) / 32)
#ifndef AWE_EXP_NO_FLUSH
#define AWE_EXP_NO_FLUSH 0
#endif
// own CCS range: per lane one row of the warp's staging area (instance-major: the row IS the run of the CCS value
// array of the lane's instance; odd row stride: conflict-free for the staging writes, lane = instance, and for the
// flush reads, lane = entry); OWN_DUMP collects outputs that are not structural entries
#define OWN_STRIDE (AWE_OWN_RS | 1)
#define OWN_DUMP (AWE_OWN_RS - 1)
#define OWN_NCVAL 8
#define OWN_MAXCST (2 * (N_X + N_Z))
#define OWN_POS_BYTES (((2 * AWE_K_COLL_MAXOWN + 15) / 16) * 16)
#define OWN_TAB_BYTES (16 * AWE_K_COLL_MAXOWNQ + OWN_POS_BYTES + ((2 * OWN_MAXCST + 15) / 16) * 16)
#define OWN_INIT(oq0, noq, ofirst, ocnt)                                                                          \
    const int oq0_ = (oq0);                                                                                        \
    (void)oq0_;                                                                                                    \
    own_init(own_tab, pr, k * N_D + j, (oq0), (noq), (ofirst) - AWE_N_KCOLL_HEAD, (ocnt), lane)
#define OWN_BEGIN(q) own_d = reinterpret_cast<const int4 *>(own_tab)[(q) - oq0_]
#define OUTO(idx, val)                                                                                            \
    do {                                                                                                           \
        /* byte offset of the output inside the lane's staging row */                                              \
        const unsigned p_ = reinterpret_cast<const unsigned short *>(own_tab + 16 * AWE_K_COLL_MAXOWNQ)[(idx)];    \
        if (BATCH_MINOR) {                                                                                         \
            if (valid && p_ != 8 * OWN_DUMP) ST_OUT(Jv + (size_t)(own_d.x + (p_ >> 3)) * (size_t)B + b, (val));    \
        } else                                                                                                     \
            STS_(own_lane + p_, (val));                                                                            \
    } while (0)
#define FLUSH_OWN(q) flush_own<BATCH_MINOR>(own_warp, own_lane, own_d, own_tab, own_cval_s, valid, n_valid, lane, Jv, b0, b, B, nnz)
#define OUT(n, val) stg[(n) * STG_STRIDE + lane] = (val)
// start of a stage-0 program: q0 = first output chunk of the group, nq = number of chunks.  The
// cooperative flush reads its destination tables from shared memory; they are fetched here with
// asynchronous copies that are older than every input copy, hence complete before the first flush.
#define GROUP_INIT(q0, nq, first)                                                                    \
    const int q0_ = (q0);                                                                            \
    (void)q0_;                                                                                       \
    do {                                                                                             \
        if (COOP) {                                                                                  \
            for (int e_ = tid; e_ < (nq) * COOP_PER; e_ += NODE_THREADS)                             \
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::                           \
                             "r"(tab_addr + e_ * 8), "l"(coop_tab + (q0) * COOP_PER + e_) : "memory"); \
            asm volatile("cp.async.commit_group;" ::: "memory");                                     \
            for (int e_ = tid; e_ < COOP_NCONST * STG_STRIDE; e_ += NODE_THREADS)                    \
                stg_u[NODE_WARPS * STG_DOUBLES + e_] = __ldg(pr.coop_const + e_ / STG_STRIDE); \
        } else                                                                                       \
            FLUSH_PREP(q0, first);                                                                   \
    } while (0)
#define FLUSH_PREP(q, first)                                                                         \
    do {                                                                                             \
        if (!COOP && !BATCH_MINOR) {                                                                 \
            LDGI_(fl_code, dest + (first) + lane % AWE_CHUNK);                                       \
            LEAF_FENCE();                                                                            \
        }                                                                                            \
    } while (0)
#define FLUSH(q, first, cnt, q_next, first_next)                                                     \
    do {                                                                                             \
        if (COOP) {                                                                                  \
            if (!AWE_EXP_NO_FLUSH) flush_coop(stg_u, tab_smem + ((q) - q0_) * COOP_PER, lane, wid, n_valid, g_tile, j_tile, n_g, nnz, bar_id); \
        } else if (BATCH_MINOR)                                                                      \
            flush_chunk_bm(stg, dest + (first), cnt, lane, valid, G, Jv, (size_t)B, (size_t)b);      \
        else {                                                                                       \
            flush_chunk(stg, fl_code, cnt, lane, n_valid, g_tile, j_tile, n_g, nnz);                 \
            if ((q_next) >= 0) FLUSH_PREP(q_next, first_next);                                       \
        }                                                                                            \
    } while (0)

// interval-cooperative flush: the N_D warps of the CTA are the N_D collocation nodes of one interval of
// the same 32 instances.  Their staged chunks together cover a (nearly) contiguous range of the CCS
// value array; lane = position in that range (destination-sorted table), warp w writes instances
// w, w + N_D, ...: every store instruction writes 32 neighbouring entries of one instance.
// barrier of the NODE_THREADS threads that own one interval (the whole CTA of the per-stage kernels, one quad
// of the fused kernel)
__device__ __forceinline__ void quad_sync(const int bar_id)
{
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NODE_THREADS) : "memory");
}

__device__ __forceinline__ void flush_coop(const double *stg_cta, const int2 *tab, const int lane, const int wid,
                                           const int n_valid, double *__restrict__ g_tile, double *__restrict__ j_tile,
                                           const int n_g, const int nnz, const int bar_id)
{
    // the destination table of the program was requested at its start (GROUP_INIT): own copies complete, then
    // the barrier makes everybody's visible
    if (AWE_POOL == 0) asm volatile("cp.async.wait_all;" ::: "memory");
    quad_sync(bar_id);
    const int n_ent = tab[COOP_PER - 1].y;        // entries of this chunk (the last table slot carries the count)
#if !defined(AWE_FLUSH_UNROLLED)
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int e = 0; e < COOP_E; ++e) {
        if (e * 32 >= n_ent) break;
        // one base pointer, 32-bit element offsets (a tile spans < 2^31 elements of g / jac values)
        const int2 t = tab[e * 32 + lane];
#ifdef AWE_EXP_CONTIG_DST
        /* ablation (wrong results): the 32 entries of a store instruction are made contiguous, to see what
           perfectly contiguous destination runs would be worth */
        const int code0 = __shfl_sync(0xffffffffu, t.x, 0);
        const int code = (t.x >= 0 && code0 >= 0) ? code0 + lane : t.x;
#else
        const int code = t.x;
#endif
        const bool is_j = code >= 0;
        double *base = is_j ? j_tile : g_tile;
        const int ld = is_j ? nnz : n_g;
        const int step = ld * N_D;
        int off = (is_j ? code : -2 - code) + wid * ld;
        const double *s = stg_cta + t.y * STG_STRIDE + wid;
        if (code != -1 && base != nullptr) {
            if (n_valid == TILE) {
                // full tile: TILE / N_D instances per warp, reads batched ahead of the stores
#if defined(AWE_FLUSH_COMPACT) && AWE_FLUSH_COMPACT >= 2
#pragma unroll 1
#else
#pragma unroll
#endif
                for (int h = 0; h < TILE / N_D; h += 4) {
                    double v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = s[(h + i) * N_D];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
#ifdef AWE_EXP_NO_STORE
                        if (v[i] == 1.2345e301) __stcs(base + (off + (h + i) * step), v[i]);   /* ablation: staging reads kept, stores (almost) never issued */
#elif defined(AWE_EXP_STORE_NEAR)
                        __stcs(base + ((off + (h + i) * step) & 0xFFFF), v[i]);   /* ablation: every store lands in the first 512 KB of the tile (L2-resident) */
#else
                        ST_OUT(base + (off + (h + i) * step), v[i]);
#endif
                    }
                }
            } else {
#pragma unroll 1
                for (int i = wid; i < n_valid; i += N_D) {
                    ST_OUT(base + off, *s);
                    off += step;
                    s += N_D;
                }
            }
        }
    }
    quad_sync(bar_id);
}

// lane = (instance group, output): output o = lane % AWE_CHUNK of the chunk, instances i = lane / AWE_CHUNK,
// + 32 / AWE_CHUNK, ...: every store instruction writes AWE_CHUNK neighbouring outputs of 32 / AWE_CHUNK
// instances
__device__ __forceinline__ void flush_chunk(double *stg, const int code_in, const int cnt, const int lane,
                                            const int n_valid, double *__restrict__ g_tile, double *__restrict__ j_tile,
                                            const int n_g, const int nnz)
{
    constexpr int GROUPS = 32 / AWE_CHUNK;
    __syncwarp();
    const int o = lane % AWE_CHUNK, i0 = lane / AWE_CHUNK;
    const int code = o < cnt ? code_in : -1;
    const bool is_j = code >= 0;
    double *base = is_j ? j_tile : g_tile;
    const int step = (is_j ? nnz : n_g) * GROUPS;
    int off = (is_j ? code : -2 - code) + i0 * (is_j ? nnz : n_g);
    const double *s = stg + o * STG_STRIDE + i0;
    if (code != -1 && base != nullptr) {
#pragma unroll 1
        for (int i = i0; i < n_valid; i += GROUPS) {
            ST_OUT(base + off, *s);
            off += step;
            s += GROUPS;
        }
    }
    __syncwarp();
}

// batch-minor layout (Jv[slot * B + b], G[row * B + b]): lane = instance, already coalesced
__device__ __forceinline__ void flush_chunk_bm(double *stg, const int *__restrict__ dest, const int cnt, const int lane,
                                               const bool valid, double *__restrict__ G, double *__restrict__ Jv,
                                               const size_t B, const size_t b)
{
    __syncwarp();
    for (int o = 0; o < cnt; ++o) {
        const int code = __ldg(dest + o);
        const double v = stg[o * STG_STRIDE + lane];
        if (!valid) continue;
        if (code >= 0) {
            if (Jv) ST_OUT(Jv + (size_t)code * B + b, v);
        } else if (code <= -2)
            ST_OUT(G + (size_t)(-2 - code) * B + b, v);
    }
    __syncwarp();
}

// tables of the own chunks oq0 .. oq0 + noq - 1 of node kj -> the warp's table area: [descriptors | positions | constants]
__device__ __forceinline__ void own_init(unsigned char *tab, const DevProblem &pr, const int kj, const int oq0, const int noq,
                                         const int ofirst, const int ocnt, const int lane)
{
    __syncwarp();
    int4 *dsc = reinterpret_cast<int4 *>(tab);
    const int4 *src = pr.own_desc + (size_t)kj * AWE_K_COLL_NOWNQ + oq0;
    int cst0 = 0, cst1 = 0;
    if (lane < noq) dsc[lane] = __ldg(src + lane);
    for (int q = 32 + lane; q < noq; q += 32) dsc[q] = __ldg(src + q);
    {
        const int4 a = __ldg(src), z = __ldg(src + noq - 1);
        cst0 = a.z, cst1 = z.z + z.w;
    }
    const unsigned char *ps = pr.own_pos + (size_t)kj * AWE_N_KCOLL_OWN + ofirst;
    unsigned short *pd = reinterpret_cast<unsigned short *>(tab + 16 * AWE_K_COLL_MAXOWNQ);
    for (int e = lane; e < ocnt; e += 32) pd[e] = (unsigned short)(8 * __ldg(ps + e));
    uchar2 *cd = reinterpret_cast<uchar2 *>(tab + 16 * AWE_K_COLL_MAXOWNQ + OWN_POS_BYTES);
    for (int e = lane; e < cst1 - cst0; e += 32) cd[e] = __ldg(pr.own_cst + cst0 + e);
    __syncwarp();
    // descriptors keep absolute constant indices: rebase to the copy
    if (lane < noq) dsc[lane].z -= cst0;
    for (int q = 32 + lane; q < noq; q += 32) dsc[q].z -= cst0;
    __syncwarp();
}

// one chunk of the own range is staged (lane = instance): constants in, then the warp writes the rows out, lane =
// entry: every store instruction writes one contiguous run of the CCS value array of one instance
template <bool BATCH_MINOR>
__device__ __forceinline__ void flush_own(const double *own_warp, const unsigned own_lane, const int4 d, const unsigned char *tab,
                                          const double *cval, const bool valid, const int n_valid, const int lane,
                                          double *__restrict__ Jv, const int b0, const int b, const int B, const int nnz)
{
    const uchar2 *cst = reinterpret_cast<const uchar2 *>(tab + 16 * AWE_K_COLL_MAXOWNQ + OWN_POS_BYTES) + d.z;
    if (BATCH_MINOR) {
        if (valid)
            for (int c = 0; c < d.w; ++c) ST_OUT(Jv + (size_t)(d.x + cst[c].x) * (size_t)B + b, cval[cst[c].y]);
        return;
    }
    for (int c = 0; c < d.w; ++c) {
        const uchar2 pc = cst[c];
        STS_(own_lane + pc.x * 8, cval[pc.y]);
    }
    __syncwarp();
    const int len = d.y;
#pragma unroll 1
    for (int p0 = 0; p0 < len; p0 += 32) {
        if (p0 + lane < len) {
            double *dst = Jv + (size_t)b0 * (size_t)nnz + d.x + p0 + lane;
            const double *src = own_warp + p0 + lane;
            if (n_valid == TILE) {
#pragma unroll 1
                for (int i = 0; i < TILE; i += 8) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = src[u * OWN_STRIDE];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        ST_OUT(dst, v[u]);
                        dst += nnz;
                    }
                    src += 8 * OWN_STRIDE;
                }
            } else {
                for (int i = 0; i < n_valid; ++i) {
                    ST_OUT(dst, *src);
                    dst += nnz, src += OWN_STRIDE;
                }
            }
        }
    }
    __syncwarp();
}

// --------------------------------------------------------------------------- gather
// grid (tiles of the chunk, n_k); b_base = first instance of the chunk
template <bool BATCH_MINOR>
__global__ void __launch_bounds__(GATHER_THREADS)
awegpu_gather_kernel(const DevProblem pr, const Scratch sc, const int B, const int b_base, const double *__restrict__ W,
                     const double *__restrict__ P, const long long ldp, const int p_off, double *__restrict__ G,
                     double *__restrict__ Jv)
{
    extern __shared__ double slab[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, k = blockIdx.y;
    const int b0 = b_base + tile * TILE;
    const int n_valid = min(TILE, B - b0);
    const int x0 = __ldg(pr.x_base + k * (N_D + 1));
    // surplus lanes of a ragged tile replicate the last instance (their results are never stored)
    for (int e = tid; e < TILE * SLAB_ALL; e += GATHER_THREADS) {
        const int u = e / SLAB_ALL, o = e - u * SLAB_ALL;
        const int b = min(b0 + u, B - 1);
        slab[u * SLAB_STRIDE + o] = __ldg(W + (size_t)b * pr.n_v + x0 + o);
    }
    __syncthreads();
    const int bl = min(b0 + lane, B - 1);
    const double *wl = W + (size_t)bl * pr.n_v;
    // ---- record: slab | theta | phi, entry-major, lane-minor
    double *rec = sc.rec + ((size_t)tile * pr.n_k + k) * REC * 32;
    for (int e = tid; e < REC * 32; e += GATHER_THREADS) {
        const int i = e >> 5;
        double v;
        if (i < SLAB_W)
            v = slab[lane * SLAB_STRIDE + i];
        else if (i < REC_PHI)
            v = __ldg(wl + __ldg(pr.theta_idx + k * N_TH + (i - REC_TH)));
        else
            v = __ldg(wl + pr.v_phi + (i - REC_PHI));
        rec[e] = v;
    }
    // ---- xdot_kj = (sum_r C[r, j+1] X_r) / h / t_f      (awebox/ocp/collocation.py:225, 239-250)
    {
        const double itf_nk = (double)pr.n_k / __ldg(wl + __ldg(pr.tf_idx + k));
        const double *s = slab + lane * SLAB_STRIDE;
        for (int j = warp; j < N_D; j += GATHER_THREADS / 32) {
            double *xd = sc.xdc + (((size_t)tile * pr.n_k + k) * N_D + j) * N_X * 32 + lane;
            for (int m = 0; m < N_X; ++m) {
                double xp = 0.0;
#pragma unroll
                for (int r = 0; r <= N_D; ++r) xp = fma(c_coeff_coll[r * (N_D + 1) + (j + 1)] * itf_nk, s[OFF_XR(r) + m], xp);
                xd[m * 32] = xp;
            }
        }
    }
    // ---- theta0, lane-contiguous (one copy per tile; a shared vector is replicated, so that the node
    //      programs address every parameter with an immediate offset)
    if (k == 0) {
        double *pt = sc.pt + (size_t)tile * pr.n_theta0 * 32;
        const double *pl = P + (size_t)bl * ldp + p_off;
        for (int e = tid; e < pr.n_theta0 * 32; e += GATHER_THREADS) pt[e] = __ldg(pl + (e >> 5));
    }
    // ---- continuity rows  x_{k+1} - sum_r D[r] X_r   (awebox/ocp/collocation.py:319-336)
    const int row_cont = __ldg(pr.row_cont + k);
    if (G == nullptr) return;                      // records only (objective path)
    if (BATCH_MINOR) {
        if (lane < n_valid)
            for (int m = warp; m < N_X; m += GATHER_THREADS / 32) {
                const double *s = slab + lane * SLAB_STRIDE;
                double xf = 0.0;
#pragma unroll
                for (int r = 0; r <= N_D; ++r) xf = fma(c_coeff_cont[r], s[OFF_XR(r) + m], xf);
                G[(size_t)(row_cont + m) * B + b0 + lane] = pr.cscale * (s[OFF_XNEXT + m] - xf);
            }
    } else {
        // lane = state component: the N_X rows of one instance are contiguous in g
        for (int u = warp; u < n_valid; u += GATHER_THREADS / 32) {
            const double *s = slab + u * SLAB_STRIDE;
            for (int m = lane; m < N_X; m += 32) {
                double xf = 0.0;
#pragma unroll
                for (int r = 0; r <= N_D; ++r) xf = fma(c_coeff_cont[r], s[OFF_XR(r) + m], xf);
                G[(size_t)(b0 + u) * pr.n_g + row_cont + m] = pr.cscale * (s[OFF_XNEXT + m] - xf);
            }
        }
    }
    // ---- constant Jacobian entries owned by this interval (continuity; periodicity on the last one):
    //      the value does not depend on the instance, thread = entry
#ifndef AWE_EXP_NO_CONST
    if (Jv) {
        const int e1 = __ldg(pr.const_begin + k + 1);
        for (int e = __ldg(pr.const_begin + k) + tid; e < e1; e += GATHER_THREADS) {
            const int slot = __ldg(pr.const_slot + e);
            const double v = __ldg(pr.const_val + e);
            if (BATCH_MINOR) {
                for (int i = 0; i < n_valid; ++i) Jv[(size_t)slot * B + b0 + i] = v;
            } else {
                for (int i = 0; i < n_valid; ++i) Jv[(size_t)(b0 + i) * pr.nnz + slot] = v;
            }
        }
    }
#endif
    // ---- periodicity rows  x_0 - x_{n_k-1, d-1}           (awebox/ocp/operation.py:245-267)
    if (k == pr.n_k - 1) {
        for (int u = warp; u < n_valid; u += GATHER_THREADS / 32) {
            const double *w0 = W + (size_t)(b0 + u) * pr.n_v + pr.x_first;
            const double *s = slab + u * SLAB_STRIDE + OFF_CX(N_D - 1);
            for (int n = lane; n < pr.n_periodic; n += 32) {
                const int m = __ldg(pr.periodic_idx + n);
                const double v = pr.cscale * (__ldg(w0 + m) - s[m]);
                if (BATCH_MINOR)
                    G[(size_t)(pr.row_periodic + n) * B + b0 + u] = v;
                else
                    G[(size_t)(b0 + u) * pr.n_g + pr.row_periodic + n] = v;
            }
        }
        // ---- affine rows: t_f bounds of single_reelout (awebox/ocp/constraints.py:147-168), initial conditions
        //      of an mpc trial (awebox/ocp/operation.py:303-326); thread = (instance, row)
        for (int e = tid; e < n_valid * pr.n_aff; e += GATHER_THREADS) {
            const int u = e / pr.n_aff, n = e - u * pr.n_aff;
            const double *w = W + (size_t)(b0 + u) * pr.n_v;
            const double *pf = P + (size_t)(b0 + u) * ldp;       // full parameter vector (checked by the host)
            double a = 0.0;
            for (int i = __ldg(pr.aff_begin + n); i < __ldg(pr.aff_begin + n + 1); ++i)
                a += __ldg(pr.aff_coef + i) * __ldg(w + __ldg(pr.aff_col + i));
            for (int i = __ldg(pr.affp_begin + n); i < __ldg(pr.affp_begin + n + 1); ++i)
                a += __ldg(pr.affp_coef + i) * __ldg(pf + __ldg(pr.affp_col + i));
            const double v = pr.cscale * (a + __ldg(pr.aff_const + n));
            const int row = __ldg(pr.aff_row + n);
            if (BATCH_MINOR)
                G[(size_t)row * B + b0 + u] = v;
            else
                G[(size_t)(b0 + u) * pr.n_g + row] = v;
        }
    }
}

// --------------------------------------------------------------------------- node programs
// Persistent, group-specialised node kernel.  One CTA per SM; a CTA is QUADS quads of NODE_WARPS warps; a quad
// works on one CTA-item at a time: warp = one node of one tile, lane = instance (collocation: the N_D nodes of one
// interval, which the cooperative flush relies on; shooting: NODE_WARPS consecutive shooting nodes).  A CTA stays on
// ONE program group for as long as that group has items left: straight-line code larger than the 32 KB instruction
// cache is limited by instruction fetch to 0.4-1.3 warp-instructions per clock and SM (profiles/r02_ubench.jsonl),
// a resident <= 27 KB program is not.  Items of a group are handed out through one atomic counter per group, so the
// share of the SMs a group gets only has to be roughly proportional to its cost (sched.cta_group, from the
// generator's weights); a CTA whose group has run dry moves on to the next group.
// Stage 0 produces the node's outputs; stages >= 1 produce exported intermediates (launched first).
#define EXPORTS_OF(SHOOT, FULL) \
    ((SHOOT) ? ((FULL) ? AWE_K_SHOOT_EXPORTS : AWE_K_SHOOT_RES_EXPORTS) : ((FULL) ? AWE_K_COLL_EXPORTS : AWE_K_COLL_RES_EXPORTS))
#define HOT_OF(SHOOT, FULL) \
    ((SHOOT) ? ((FULL) ? AWE_K_SHOOT_HOT : AWE_K_SHOOT_RES_HOT) : ((FULL) ? AWE_K_COLL_HOT : AWE_K_COLL_RES_HOT))
#define SMEM_PER_SM (227 * 1024)
#define MAX_GROUPS 64

#define MAX_CTAS 192
struct Sched {
    int *counters;            // [MAX_GROUPS] next CTA-item of every group of this launch (zeroed before the launch)
    // group a CTA starts with.  By value: a kernel parameter indexed by blockIdx is a uniform value to ptxas, so that
    // the switch over the program groups is a uniform branch and the warp barriers inside the programs (LEAF_FENCE)
    // cost nothing; a value loaded from global or shared memory is not, and every LEAF_FENCE becomes a
    // WARPSYNC.COLLECTIVE that waits for the outstanding loads
    unsigned char cta_group[MAX_CTAS];
    int lead;                 // CTA-items a group may run ahead of the slowest one before its CTAs move over
};

// shared memory of a quad: [spill slots | tables of the own chunks (per warp) | U], where U is either the chunk
// staging of the cooperative / per-warp flush + replicated constants + destination tables, or the instance-major
// rows of the own CCS range (a program group uses one of the two layouts)
template <bool SHOOT, bool FULL>
__host__ __device__ static constexpr size_t own_tab_doubles()
{
    return (!SHOOT && FULL) ? ((size_t)NODE_WARPS * OWN_TAB_BYTES + 7) / 8 : 0;
}
template <bool SHOOT, bool FULL>
__host__ __device__ static constexpr size_t union_doubles()
{
    const size_t head = (size_t)NODE_WARPS * STG_DOUBLES +
                        ((!SHOOT && FULL) ? (size_t)COOP_NCONST * STG_STRIDE + (size_t)AWE_K_COLL_MAXQ * COOP_PER : 0);
    const size_t own = (!SHOOT && FULL) ? (size_t)NODE_THREADS * OWN_STRIDE : 0;
    return head > own ? head : own;
}
template <bool SHOOT, bool FULL>
__host__ __device__ static constexpr size_t node_smem_bytes()
{
    return sizeof(double) * ((size_t)HOT_OF(SHOOT, FULL) * NODE_THREADS + own_tab_doubles<SHOOT, FULL>() + union_doubles<SHOOT, FULL>());
}
// quads per CTA: what the shared memory of an SM holds, at most AWE_NODE_QUADS (3 quads = 12 warps leave every
// thread 168 registers)
#ifndef AWE_NODE_QUADS
#define AWE_NODE_QUADS 3
#endif
template <bool SHOOT, bool FULL>
__host__ __device__ static constexpr int node_quads()
{
    return (SMEM_PER_SM - 64) / node_smem_bytes<SHOOT, FULL>() >= AWE_NODE_QUADS
               ? AWE_NODE_QUADS
               : ((SMEM_PER_SM - 64) / node_smem_bytes<SHOOT, FULL>() >= 1 ? (int)((SMEM_PER_SM - 64) / node_smem_bytes<SHOOT, FULL>()) : 1);
}
template <bool SHOOT, bool FULL, int STAGE>
__host__ __device__ static constexpr int node_groups()
{
    return SHOOT ? (FULL ? (STAGE == 0 ? AWE_K_SHOOT_S0_GROUPS : STAGE == 1 ? AWE_K_SHOOT_S1_GROUPS : AWE_K_SHOOT_S2_GROUPS)
                         : (STAGE == 0 ? AWE_K_SHOOT_RES_S0_GROUPS
                                       : STAGE == 1 ? AWE_K_SHOOT_RES_S1_GROUPS : AWE_K_SHOOT_RES_S2_GROUPS))
                 : (FULL ? (STAGE == 0 ? AWE_K_COLL_S0_GROUPS : STAGE == 1 ? AWE_K_COLL_S1_GROUPS : AWE_K_COLL_S2_GROUPS)
                         : (STAGE == 0 ? AWE_K_COLL_RES_S0_GROUPS
                                       : STAGE == 1 ? AWE_K_COLL_RES_S1_GROUPS : AWE_K_COLL_RES_S2_GROUPS));
}

template <bool SHOOT, bool FULL, int STAGE, bool BATCH_MINOR>
__global__ void __launch_bounds__(NODE_THREADS * node_quads<SHOOT, FULL>(), 1)
awegpu_node_kernel(const DevProblem pr, const Scratch sc, const int B, const int b_base, const int n_items,
                   const double *__restrict__ p_shared, const Sched sched, double *__restrict__ G, double *__restrict__ Jv)
{
    constexpr int HOT = HOT_OF(SHOOT, FULL);
    constexpr int NEX = EXPORTS_OF(SHOOT, FULL);
    constexpr int QUADS = node_quads<SHOOT, FULL>();
    constexpr int N_GROUPS = node_groups<SHOOT, FULL, STAGE>();
    constexpr size_t QUAD_DOUBLES = node_smem_bytes<SHOOT, FULL>() / sizeof(double);
    // the cooperative flush needs the quad to be exactly the N_D nodes of one interval
#ifdef AWE_EXP_NO_COOP
    constexpr bool COOP = false;          /* ablation: per-warp flush for the collocation nodes as well */
#else
    constexpr bool COOP = !SHOOT && FULL && STAGE == 0 && !BATCH_MINOR && N_D == NODE_WARPS;
#endif
    extern __shared__ double smem_all[];
    __shared__ int s_next[QUADS][2], s_group[QUADS][2];
    const int quad = threadIdx.x / NODE_THREADS;
    const int tid = threadIdx.x - quad * NODE_THREADS, lane = tid & 31, wid = tid >> 5;
    const int bar_id = 1 + quad;
    double *smem = smem_all + (size_t)quad * QUAD_DOUBLES;
    double *stg_u = smem + (size_t)HOT * NODE_THREADS + own_tab_doubles<SHOOT, FULL>();
    double *stg = stg_u + wid * STG_DOUBLES;
    const unsigned scr_addr = (unsigned)__cvta_generic_to_shared(smem + tid);
    const int2 *tab_smem = reinterpret_cast<const int2 *>(stg_u + NODE_WARPS * STG_DOUBLES + COOP_NCONST * STG_STRIDE);
    const unsigned tab_addr = (unsigned)__cvta_generic_to_shared(tab_smem);
    // own CCS range (collocation nodes, stage 0): the warp's table area and the lane's staging row
    unsigned char *own_tab = reinterpret_cast<unsigned char *>(smem + (size_t)HOT * NODE_THREADS) + (size_t)wid * OWN_TAB_BYTES;
    const double *own_warp = stg_u + (size_t)wid * 32 * OWN_STRIDE;
    const unsigned own_lane = (unsigned)__cvta_generic_to_shared(stg_u + (size_t)tid * OWN_STRIDE);
    int4 own_d = make_int4(0, 0, 0, 0);
    __shared__ double own_cval_s[OWN_NCVAL];
    if (!SHOOT && FULL && STAGE == 0) {
        if (threadIdx.x < OWN_NCVAL) own_cval_s[threadIdx.x] = __ldg(pr.own_cval + threadIdx.x);
        __syncthreads();
    }
    (void)own_tab, (void)own_lane, (void)own_warp, (void)own_d;
    const int n_g = pr.n_g;
    const int nnz = (int)pr.nnz;
    const int n_cta_items = (n_items + NODE_WARPS - 1) / NODE_WARPS;
    constexpr bool EX_RW = false;
    (void)p_shared, (void)stg, (void)scr_addr, (void)tab_smem, (void)tab_addr;
    int group = sched.cta_group[blockIdx.x];
    int it = 0;
    // A CTA leaves its group for the one that lags most when its own group is more than `lead` CTA-items ahead of
    // it: the groups of one tile then run within a short time of each other, so that the records and exports they
    // share are read from L2 and the sectors they share in the output are completed in L2 (a static split lets
    // the groups drift apart by dozens of tiles).  Also the end game: a group that has run dry counts as "ahead".
    const int lead = sched.lead;
    // (Choosing the group one item early, to hide the latency of the counter reads and of the atomic, was measured
    // 5 - 50 % slower: with stale counters the CTAs flock to the same lagging group and the groups drift apart.)
#pragma unroll 1
    while (true) {
        // next CTA-item: one atomic per quad, handed to the other threads through shared memory (two slots: a
        // thread may be one iteration ahead of the slowest thread of its quad, never two)
        if (wid == 0) {
            int g = group;
            int mine = 0x7fffffff, best = 0x7fffffff, best_g = 0;
            for (int g0 = 0; g0 < N_GROUPS; g0 += 32) {
                const int gg = g0 + lane;
                const int c = gg < N_GROUPS ? *(volatile const int *)(sched.counters + gg) : 0x7fffffff;
                const int m = __reduce_min_sync(0xffffffffu, c);
                if (m < best) best = m, best_g = g0 + __ffs(__ballot_sync(0xffffffffu, c == m)) - 1;
                const int cm = __shfl_sync(0xffffffffu, c, g & 31);
                if ((g >> 5) == (g0 >> 5)) mine = cm;
            }
            if (mine >= n_cta_items || mine - best > lead) g = best_g;
            if (lane == 0) {
                s_next[quad][it & 1] = atomicAdd(sched.counters + g, 1);
                s_group[quad][it & 1] = g;
            }
        }
        quad_sync(bar_id);
        const int cta_item = s_next[quad][it & 1];
        {
            // votes: warp-uniform values by construction (see Sched)
            const int g = s_group[quad][it & 1];
            int gu = 0;
#pragma unroll
            for (int bit = 0; bit < 6; ++bit) gu |= __any_sync(0xffffffffu, (g >> bit) & 1) ? (1 << bit) : 0;
            group = gu;
        }
        ++it;
        if (__all_sync(0xffffffffu, cta_item >= n_cta_items)) break;

        {
            const int item_raw = cta_item * NODE_WARPS + wid;
            const bool active = __all_sync(0xffffffffu, item_raw < n_items);   // collocation (N_D == NODE_WARPS): whole quads are active
            const int item = active ? item_raw : n_items - 1;
            int tile, k, j;
            if (SHOOT) {
                j = -1, k = item % pr.n_k, tile = item / pr.n_k;
            } else {
                j = item % N_D, k = (item / N_D) % pr.n_k, tile = item / (N_D * pr.n_k);
            }
            const int j1 = j + 1;
            const int b0 = b_base + tile * TILE;
            const int b = b0 + lane;
            const bool valid = b < B;
            const int n_valid = min(TILE, B - b0);
            const double *w_rec = sc.rec + ((size_t)tile * pr.n_k + k) * REC * 32 + lane;
            const double *w_own = SHOOT ? w_rec + OFF_X0 * 32 : w_rec + OFF_CX(j) * 32;
            const double *w_z = SHOOT ? w_rec + OFF_Z * 32 : w_own + N_X * 32;
            const double *w_xd = SHOOT ? w_rec + OFF_XDOT * 32
                                       : sc.xdc + (((size_t)tile * pr.n_k + k) * N_D + j) * N_X * 32 + lane;
            // theta0: the lane-contiguous copy the gather kernel made (per-instance vectors, or the shared one replicated)
            const double *w_pt = sc.pt + (size_t)tile * pr.n_theta0 * 32 + lane;
            double *w_ex = (SHOOT ? sc.exs : sc.exc) + (size_t)item * (NEX > 0 ? NEX : 1) * 32 + lane;
            // siblings of the interval: record offsets of their state blocks, offsets of their contracted derivatives
            // relative to the node's own, their block indices
            int sib_off[N_D > 1 ? N_D - 1 : 1], sib_xd[N_D > 1 ? N_D - 1 : 1], sib_j1[N_D > 1 ? N_D - 1 : 1];
#pragma unroll
            for (int dl = 1; dl < N_D; ++dl) {
                const int sj = SHOOT ? 0 : (j + dl) % N_D;
                sib_off[dl - 1] = OFF_CX(sj) * 32, sib_xd[dl - 1] = (sj - j) * (N_X * 32), sib_j1[dl - 1] = sj + 1;
            }
            (void)sib_off, (void)sib_xd, (void)sib_j1;
            const double itf = 1.0 / __ldg(w_rec + (REC_TH + (AWE_TF_LOCAL - (2 * N_X + N_U + N_Z))) * 32);
            const double itf_nk = itf * (double)pr.n_k;
            const int *dest = SHOOT ? pr.shoot_dest + (size_t)k * pr.n_shoot_out
                                    : pr.coll_dest + (size_t)(k * N_D + j) * pr.n_coll_out;
            double *g_tile = G + (size_t)b0 * n_g;
            double *j_tile = Jv ? Jv + (size_t)b0 * (size_t)nnz : nullptr;
            const int2 *coop_tab = pr.coop_tab + (size_t)k * AWE_K_COLL_NCHUNKS * COOP_PER;
            int fl_code = -1;
            (void)itf_nk, (void)w_ex, (void)valid, (void)n_valid, (void)w_z, (void)w_pt, (void)w_xd, (void)w_own, (void)j1, (void)b;
            (void)dest, (void)g_tile, (void)j_tile, (void)coop_tab, (void)fl_code;
            if (active) {
                if (SHOOT && FULL) {
                    if (STAGE == 0) {
                        switch (group) {
#include "k_shoot_s0.inc"
                        }
                    } else if (STAGE == 1) {
                        switch (group) {
#include "k_shoot_s1.inc"
                        }
                    } else {
                        switch (group) {
#include "k_shoot_s2.inc"
                        }
                    }
                } else if (SHOOT && !FULL) {
                    if (STAGE == 0) {
                        switch (group) {
#include "k_shoot_res_s0.inc"
                        }
                    } else if (STAGE == 1) {
                        switch (group) {
#include "k_shoot_res_s1.inc"
                        }
                    } else {
                        switch (group) {
#include "k_shoot_res_s2.inc"
                        }
                    }
                } else if (!SHOOT && FULL) {
                    if (STAGE == 0) {
                        switch (group) {
#include "k_coll_s0.inc"
                        }
                    } else if (STAGE == 1) {
                        switch (group) {
#include "k_coll_s1.inc"
                        }
                    } else {
                        switch (group) {
#include "k_coll_s2.inc"
                        }
                    }
                } else {
                    if (STAGE == 0) {
                        switch (group) {
#include "k_coll_res_s0.inc"
                        }
                    } else if (STAGE == 1) {
                        switch (group) {
#include "k_coll_res_s1.inc"
                        }
                    } else {
                        switch (group) {
#include "k_coll_res_s2.inc"
                        }
                    }
                }
            }
        }
    }
}

// --------------------------------------------------------------------------- launch plumbing (per translation unit)
struct NodeLaunch {
    int n_ctas;               // SMs of the device
    cudaStream_t st;
    DevProblem dp;
    Scratch sc;
    int B, b_base, n_items;
    const double *p_sh;
    int *counters;            // [6][MAX_GROUPS]: one block per (collocation | shooting, stage), zeroed by the caller
    const unsigned char *cta_group;     // [6][n_ctas] (host)
    int lead;
    double *G, *Jv;
};

template <bool S, bool F, int T>
static void launch_node_impl(const bool bm, const NodeLaunch &a)
{
    constexpr int kind = (S ? 3 : 0) + T;
    constexpr int quads = node_quads<S, F>();
    const int n_cta_items = (a.n_items + NODE_WARPS - 1) / NODE_WARPS;
    // no more CTAs than there are quads' worth of work
    const long long work = (long long)n_cta_items * node_groups<S, F, T>();
    const int grid = (int)std::min<long long>(a.n_ctas, (work + quads - 1) / quads);
    Sched sched;
    sched.counters = a.counters + kind * MAX_GROUPS;
    sched.lead = a.lead;
    // a smaller grid samples the table evenly
    for (int c = 0; c < grid; ++c) sched.cta_group[c] = a.cta_group[(size_t)kind * a.n_ctas + (size_t)c * a.n_ctas / grid];
    const size_t smem = quads * node_smem_bytes<S, F>();
    // the output layout only matters to the output stage: pre-stages have one instantiation
    if (bm && T == 0 && AWE_BATCH_MINOR)
        awegpu_node_kernel<S, F, T, (T == 0) && AWE_BATCH_MINOR><<<grid, NODE_THREADS * quads, smem, a.st>>>(
            a.dp, a.sc, a.B, a.b_base, a.n_items, a.p_sh, sched, a.G, a.Jv);
    else
        awegpu_node_kernel<S, F, T, false><<<grid, NODE_THREADS * quads, smem, a.st>>>(
            a.dp, a.sc, a.B, a.b_base, a.n_items, a.p_sh, sched, a.G, a.Jv);
}

template <bool S, bool F, int T>
static cudaError_t attr_node_impl()
{
    const int bytes = (int)(node_quads<S, F>() * node_smem_bytes<S, F>());
    cudaError_t e = cudaFuncSetAttribute(awegpu_node_kernel<S, F, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && T == 0 && AWE_BATCH_MINOR)
        e = cudaFuncSetAttribute(awegpu_node_kernel<S, F, T, (T == 0) && AWE_BATCH_MINOR>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}

// every unit has its own copy of the __constant__ coefficient tables (no relocatable device code)
static cudaError_t set_coeff_here(const double *coeff_coll, const double *coeff_cont)
{
    cudaError_t e = cudaMemcpyToSymbol(c_coeff_coll, coeff_coll, sizeof(double) * (N_D + 1) * (N_D + 1));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_coeff_cont, coeff_cont, sizeof(double) * (N_D + 1));
    return e;
}

#define DECL_TU(n)                                                              \
    void awe_launch_tu##n(int stage, bool bm, const NodeLaunch &a);             \
    cudaError_t awe_init_tu##n(const double *coeff_coll, const double *coeff_cont);
DECL_TU(1)
DECL_TU(2)
DECL_TU(3)
DECL_TU(4)
#undef DECL_TU

#if TU_OWNS(1)
void awe_launch_tu1(int, bool bm, const NodeLaunch &a) { launch_node_impl<false, true, 0>(bm, a); }
cudaError_t awe_init_tu1(const double *cc, const double *cf)
{
    cudaError_t e = set_coeff_here(cc, cf);
    return e != cudaSuccess ? e : attr_node_impl<false, true, 0>();
}
#endif
#if TU_OWNS(2)
void awe_launch_tu2(int stage, bool bm, const NodeLaunch &a)
{
    if (stage == 1) launch_node_impl<false, true, 1>(bm, a);
    else launch_node_impl<false, true, 2>(bm, a);
}
cudaError_t awe_init_tu2(const double *cc, const double *cf)
{
    cudaError_t e = set_coeff_here(cc, cf);
    if (e == cudaSuccess) e = attr_node_impl<false, true, 1>();
    return e != cudaSuccess ? e : attr_node_impl<false, true, 2>();
}
#endif
#if TU_OWNS(3)
void awe_launch_tu3(int, bool bm, const NodeLaunch &a) { launch_node_impl<true, true, 0>(bm, a); }
cudaError_t awe_init_tu3(const double *cc, const double *cf)
{
    cudaError_t e = set_coeff_here(cc, cf);
    return e != cudaSuccess ? e : attr_node_impl<true, true, 0>();
}
#endif
#if TU_OWNS(4)
void awe_launch_tu4(int stage, bool bm, const NodeLaunch &a)
{
    if (stage == 1) launch_node_impl<true, true, 1>(bm, a);
    else launch_node_impl<true, true, 2>(bm, a);
}
cudaError_t awe_init_tu4(const double *cc, const double *cf)
{
    cudaError_t e = set_coeff_here(cc, cf);
    if (e == cudaSuccess) e = attr_node_impl<true, true, 1>();
    return e != cudaSuccess ? e : attr_node_impl<true, true, 2>();
}
#endif

#if TU_OWNS(0)
// full-Jacobian kernels live in the other units, residual-only ones here
template <bool S, bool F, int T>
static void launch_node(const bool bm, const NodeLaunch &a)
{
    if (!F) launch_node_impl<S, false, T>(bm, a);
    else if (!S && T == 0) awe_launch_tu1(T, bm, a);
    else if (!S) awe_launch_tu2(T, bm, a);
    else if (T == 0) awe_launch_tu3(T, bm, a);
    else awe_launch_tu4(T, bm, a);
}
#endif



#if TU_OWNS(0)
// --------------------------------------------------------------------------- objective f(w, p) and its gradient
// (awebox_b200/ocp/objective.py restates awebox/ocp/objective.py)
// awegpu_obj_kernel: CTA = one instance, no side buffers.  The regularisation terms are element-wise in V: warp = one
// interval, lanes on consecutive entries (coalesced reads of w and of the reference straight from the instance-major
// rows, coalesced plain stores of the interval block of grad f), including xdot = C X / (h t_f) and its chain into all
// X_{k,r} and t_f.  The integrands that come from the model (power, side-slip angles; generated: node_cost.inc) need a
// node's variables together: the interval blocks of OBJ_LSQ_WARPS intervals stay in shared memory, and one warp of the
// CTA (taking turns) evaluates their d nodes, thread = node, adding the few gradient entries as L2 reductions.
// Sums go to per-instance accumulators; a second kernel (thread = instance) combines them:
// f = psi * tracking + (1 - psi) * power + general + homotopy.
#define OBJ_NACC 8           /* reg[5], beta, energy, spare */
struct ObjParams {
    const int *cat;          // [N_VAR] cost category of every entry of the node column (-1: none)
    double norm[6];          // normalisation of the five categories and of the beta cost
    double quad[N_D];        // quadrature weights of the collocation nodes
    int n_tf, tf_idx[2];     // time period T = sum tf_coef * V[tf_idx]
    double tf_coef[2];
    int p_weights, p_cost, psi_idx, n_phi_cost;
    int phi_cost[16];        // index into the cost block of every homotopy parameter (order of V.phi)
    int c_power, c_tf, c_beta, c_cat[5];
};

#define OBJ_LSQ_WARPS 8
#define OBJ_N_SMALL (N_U + N_D * N_Z + N_TH)
#define OBJ_NXC ((N_X + 31) / 32)
#define OBJ_NSC ((OBJ_N_SMALL + 31) / 32)
// CTA = one instance, warp = every OBJ_LSQ_WARPS-th interval, lane = state index m (all X_{k,r}[m] of the interval) and
// one entry of the short segments (u_k; z_kj; theta).  Everything that does not depend on the interval (category,
// weight, factors) is set up once per lane; all loads of an interval are independent of each other: one round trip to
// memory per interval and warp.  Sums of squares are kept per lane and weighted once at the end.
// entry i of the column of collocation node j (x, xdot, u, z, theta), read from the staged interval block
__device__ __forceinline__ double obj_vin(const int i, const double *s, const int j, const double itf_nk, const double *Wb,
                                          const int *th_idx)
{
    if (i < N_X) return s[OFF_CX(j) + i];
    if (i < 2 * N_X) {
        double xd = 0.0;
#pragma unroll
        for (int r = 0; r <= N_D; ++r) xd = fma(c_coeff_coll[r * (N_D + 1) + (j + 1)], s[OFF_XR(r) + (i - N_X)], xd);
        return xd * itf_nk;
    }
    if (i < 2 * N_X + N_U) return s[OFF_U + (i - 2 * N_X)];
    if (i < 2 * N_X + N_U + N_Z) return s[OFF_CZ(j) + (i - 2 * N_X - N_U)];
    return __ldg(Wb + __ldg(th_idx + (i - 2 * N_X - N_U - N_Z)));
}

// intervals staged at a time: all of them when they fit into a third of the SM's shared memory (three CTAs per SM)
static int obj_stage_k(const int n_k)
{
    int s = (int)((72 * 1024) / (sizeof(double) * OFF_XNEXT));
    s -= s % OBJ_LSQ_WARPS;
    if (s < OBJ_LSQ_WARPS) s = OBJ_LSQ_WARPS;
    const int all = (n_k + OBJ_LSQ_WARPS - 1) / OBJ_LSQ_WARPS * OBJ_LSQ_WARPS;
    return s < all ? s : all;
}
static size_t obj_smem_bytes(const int n_k) { return sizeof(double) * (size_t)obj_stage_k(n_k) * OFF_XNEXT; }
static_assert(sizeof(double) * OBJ_LSQ_WARPS * OFF_XNEXT <= 200 * 1024, "objective kernel: interval blocks do not fit in shared memory");

__global__ void __launch_bounds__(32 * OBJ_LSQ_WARPS, 3)
awegpu_obj_kernel(const DevProblem pr, const ObjParams ob, const int B, const double *__restrict__ W,
                  const double *__restrict__ P, const long long ldp, const int p_theta0, const int stage_k,
                  const int *__restrict__ rest, const int n_rest, double *__restrict__ F, double *__restrict__ grad,
                  double *__restrict__ comp)
{
    extern __shared__ double obj_blk[];       // [stage_k][OFF_XNEXT]: interval blocks of w
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.x;
    const double *Wb = W + (size_t)b * pr.n_v, *Pb = P + (size_t)b * ldp;
    const double *Pw = Pb + ob.p_weights;
    double *gb = grad + (size_t)b * pr.n_v;
    const double psi = __ldg(Wb + ob.psi_idx);
    double cat_f[5], q_sum = 0.0;
#pragma unroll
    for (int c = 0; c < 5; ++c) cat_f[c] = __ldg(Pb + ob.p_cost + ob.c_cat[c]) / ob.norm[c];
#pragma unroll
    for (int j = 0; j < N_D; ++j) q_sum += ob.quad[j];
    // weight * category factor of an entry of the node column with category c (0: no cost)
    auto factor = [&](const int c, const double w) -> double {
        double cf = (c == 0) ? cat_f[0] : 0.0;
#pragma unroll
        for (int cc = 1; cc < 5; ++cc) cf = (c == cc) ? cat_f[cc] : cf;
        return w * cf;
    };
    // ---- per lane, the same for every interval
    int cx[OBJ_NXC], cd[OBJ_NXC];
    double fx[OBJ_NXC], fd[OBJ_NXC], hx[OBJ_NXC], hd[OBJ_NXC], sx[OBJ_NXC], sd[OBJ_NXC];
#pragma unroll
    for (int ch = 0; ch < OBJ_NXC; ++ch) {
        const int m = min(lane + 32 * ch, N_X - 1);
        cx[ch] = __ldg(ob.cat + m), cd[ch] = __ldg(ob.cat + N_X + m);
        fx[ch] = factor(cx[ch], __ldg(Pw + m)), fd[ch] = factor(cd[ch], __ldg(Pw + N_X + m));
        hx[ch] = 2.0 * fx[ch] * (cx[ch] == 0 ? psi : 1.0), hd[ch] = 2.0 * fd[ch] * (cd[ch] == 0 ? psi : 1.0);
        sx[ch] = sd[ch] = 0.0;
    }
    // short segments: controls (the same u_k at every node of the interval: the quadrature weights add up), algebraic
    // variables of the d collocation nodes, theta (shared with the other intervals: L2 reductions)
    int cs[OBJ_NSC], so[OBJ_NSC];          // category; offset from x_k (u, z) or -1 - t (theta entry t)
    double fs[OBJ_NSC], hs[OBJ_NSC], qs[OBJ_NSC], ss[OBJ_NSC];
#pragma unroll
    for (int ch = 0; ch < OBJ_NSC; ++ch) {
        const int e = min(lane + 32 * ch, OBJ_N_SMALL - 1);
        int i;
        if (e < N_U) i = 2 * N_X + e, so[ch] = OFF_U + e, qs[ch] = q_sum;
        else if (e < N_U + N_D * N_Z) {
            const int j = (e - N_U) / N_Z, m = (e - N_U) - j * N_Z;
            i = 2 * N_X + N_U + m, so[ch] = OFF_CZ(j) + m, qs[ch] = ob.quad[j];
        } else
            i = 2 * N_X + N_U + N_Z + (e - N_U - N_D * N_Z), so[ch] = -1 - (e - N_U - N_D * N_Z), qs[ch] = q_sum;
        cs[ch] = __ldg(ob.cat + i);
        fs[ch] = factor(cs[ch], __ldg(Pw + i));
        hs[ch] = 2.0 * fs[ch] * (cs[ch] == 0 ? psi : 1.0);
        ss[ch] = 0.0;
    }
    // entries of grad f outside the interval blocks (theta, phi, the last state ...) only receive reductions: zero first
    for (int i = threadIdx.x; i < n_rest; i += 32 * OBJ_LSQ_WARPS) gb[__ldg(rest + i)] = 0.0;
    __syncthreads();
    double e_sum = 0.0, b_sum = 0.0;          // integrands: energy and squared side-slip angles of the nodes of this lane
    const double c_power = __ldg(Pb + ob.p_cost + ob.c_power);
    const double f_beta = __ldg(Pb + ob.p_cost + ob.c_beta) / ob.norm[5];
    double T = 0.0;
    for (int i = 0; i < ob.n_tf; ++i) T += ob.tf_coef[i] * __ldg(Wb + ob.tf_idx[i]);
    for (int k_base = 0; k_base < pr.n_k; k_base += stage_k) {
    const int k_end = min(pr.n_k, k_base + stage_k);
    for (int k = k_base + wid; k < k_end; k += OBJ_LSQ_WARPS) {
        double *blk = obj_blk + (k - k_base) * OFF_XNEXT;
        const int x0 = __ldg(pr.x_base + k * (N_D + 1));
        const int tf_v = __ldg(pr.tf_idx + k);
        // ---- loads: the interval block of w through shared memory, the references straight from p
        {
            double t[(OFF_XNEXT + 31) / 32];
#pragma unroll
            for (int it = 0; it < (OFF_XNEXT + 31) / 32; ++it) t[it] = __ldg(Wb + x0 + min(lane + 32 * it, OFF_XNEXT - 1));
#pragma unroll
            for (int it = 0; it < (OFF_XNEXT + 31) / 32; ++it)
                if (lane + 32 * it < OFF_XNEXT) blk[lane + 32 * it] = t[it];
        }
        int si[OBJ_NSC];
        double sw[OBJ_NSC], sr[OBJ_NSC], X[OBJ_NXC][N_D + 1], R[OBJ_NXC][N_D + 1];
#pragma unroll
        for (int ch = 0; ch < OBJ_NSC; ++ch) {
            si[ch] = so[ch] >= 0 ? x0 + so[ch] : __ldg(pr.theta_idx + k * N_TH + (-1 - so[ch]));
            sr[ch] = __ldg(Pb + si[ch]);
            if (so[ch] < 0) sw[ch] = __ldg(Wb + si[ch]);
        }
#pragma unroll
        for (int ch = 0; ch < OBJ_NXC; ++ch) {
            const int m = min(lane + 32 * ch, N_X - 1);
#pragma unroll
            for (int r = 1; r <= N_D; ++r) R[ch][r] = __ldg(Pb + x0 + OFF_XR(r) + m);
        }
        const double tf = __ldg(Wb + tf_v);
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < OBJ_NSC; ++ch)
            if (so[ch] >= 0) sw[ch] = blk[so[ch]];
#pragma unroll
        for (int ch = 0; ch < OBJ_NXC; ++ch) {
            const int m = min(lane + 32 * ch, N_X - 1);
#pragma unroll
            for (int r = 0; r <= N_D; ++r) X[ch][r] = blk[OFF_XR(r) + m];
        }
        const double itf = 1.0 / tf, itf_nk = (double)pr.n_k * itf;
        double tf_g = 0.0;
        // ---- differential states: X_{k,0} = x_k and the collocation states X_{k,r}
#pragma unroll
        for (int ch = 0; ch < OBJ_NXC; ++ch) {
            const int m = lane + 32 * ch;
            double g[N_D + 1];
            g[0] = 0.0;
#pragma unroll
            for (int r = 1; r <= N_D; ++r) {
                const double qd = ob.quad[r - 1] * (X[ch][r] - R[ch][r]);
                sx[ch] = fma(qd, X[ch][r] - R[ch][r], sx[ch]);
                g[r] = hx[ch] * qd;
            }
            // xdot_kj = (sum_r C[r, j+1] X_r) / h / t_f, reference 0;  d xdot_kj / d X_r = C[r, j+1] / h / t_f,
            // d xdot_kj / d t_f = -xdot_kj / t_f
            double tf_l = 0.0;
#pragma unroll
            for (int j = 0; j < N_D; ++j) {
                double xd = 0.0;
#pragma unroll
                for (int r = 0; r <= N_D; ++r) xd = fma(c_coeff_coll[r * (N_D + 1) + (j + 1)], X[ch][r], xd);
                xd *= itf_nk;
                const double qd = ob.quad[j] * xd;
                sd[ch] = fma(qd, xd, sd[ch]);
                const double gxd = hd[ch] * qd, gs = gxd * itf_nk;
#pragma unroll
                for (int r = 0; r <= N_D; ++r) g[r] = fma(gs, c_coeff_coll[r * (N_D + 1) + (j + 1)], g[r]);
                tf_l = fma(gxd, xd, tf_l);
            }
            if (m < N_X) {
                tf_g -= tf_l * itf;
#pragma unroll
                for (int r = 0; r <= N_D; ++r) gb[x0 + OFF_XR(r) + m] = g[r];
                gb[x0 + OFF_XDOT + m] = 0.0;          // xdot_k, z_k: variables of the shooting node, no cost
            } else
                sx[ch] = sd[ch] = 0.0;       // a clamped lane: only lanes of the last chunk, which stay clamped
        }
        for (int m = lane; m < N_Z; m += 32) gb[x0 + OFF_Z + m] = 0.0;
        // ---- short segments
#pragma unroll
        for (int ch = 0; ch < OBJ_NSC; ++ch) {
            const double diff = sw[ch] - sr[ch], qd = qs[ch] * diff;
            if (lane + 32 * ch < OBJ_N_SMALL) {
                ss[ch] = fma(qd, diff, ss[ch]);
                const double gv = hs[ch] * qd;
                if (so[ch] >= 0) gb[si[ch]] = gv;
                else if (gv != 0.0) atomicAdd(gb + si[ch], gv);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tf_g += __shfl_xor_sync(0xffffffffu, tf_g, o);
        if (lane == 0 && tf_g != 0.0) atomicAdd(gb + tf_v, tf_g);
    }
        __syncthreads();      // the blocks of these intervals are staged and their part of grad f is stored
        // ---- integrands from the model at the d nodes of the staged intervals: thread = node
        {
            const int n_nodes = (k_end - k_base) * N_D;
            for (int n = threadIdx.x; n < n_nodes; n += 32 * OBJ_LSQ_WARPS) {
                const int kk = n / N_D, j = n - kk * N_D, kn = k_base + kk;
                const double *sb = obj_blk + kk * OFF_XNEXT;
                const int x0 = __ldg(pr.x_base + kn * (N_D + 1));
                const int tf_v = __ldg(pr.tf_idx + kn);
                const int *th_idx = pr.theta_idx + kn * N_TH;
                const double tf = __ldg(Wb + tf_v);
                const double itf = 1.0 / tf, itf_nk = (double)pr.n_k * itf;
                const double q = ob.quad[j];
                // chain a derivative with respect to entry i of the node column into grad f
                auto scatter = [&](const int i, const double gv) {
                    if (gv == 0.0) return;
                    if (i < N_X)
                        atomicAdd(gb + x0 + OFF_CX(j) + i, gv);
                    else if (i < 2 * N_X) {
                        const int m = i - N_X;
                        for (int r = 0; r <= N_D; ++r)
                            atomicAdd(gb + x0 + OFF_XR(r) + m, gv * c_coeff_coll[r * (N_D + 1) + (j + 1)] * itf_nk);
                        atomicAdd(gb + tf_v, -gv * obj_vin(i, sb, j, itf_nk, Wb, th_idx) * itf);
                    } else if (i < 2 * N_X + N_U)
                        atomicAdd(gb + x0 + OFF_U + (i - 2 * N_X), gv);
                    else if (i < 2 * N_X + N_U + N_Z)
                        atomicAdd(gb + x0 + OFF_CZ(j) + (i - 2 * N_X - N_U), gv);
                    else
                        atomicAdd(gb + __ldg(th_idx + (i - 2 * N_X - N_U - N_Z)), gv);
                };
                double out[AWE_COST_N_OUT];
                {
#define VIN(i) obj_vin((i), sb, j, itf_nk, Wb, th_idx)
#define PIN(i) ((i) < AWE_N_THETA0 ? __ldg(Pb + p_theta0 + (i)) : __ldg(Wb + pr.v_phi + ((i) - AWE_N_THETA0)))
#pragma push_macro("OUT")
#undef OUT
#define OUT(n, val) out[n] = (val)
#include "node_cost.inc"
#pragma pop_macro("OUT")
#undef VIN
#undef PIN
                }
                int pos = 0;
#pragma unroll
                for (int fn = 0; fn < AWE_COST_N_FUN; ++fn) {
                    const int first = awe_cost_layout[pos], cnt = awe_cost_layout[pos + 1];
                    const double val = out[first];
                    if (fn == 0) {
                        // energy += h q p;  power cost = -c_power E / T, weighted by (1 - psi)
                        const double h = tf / (double)pr.n_k;
                        e_sum = fma(h * q, val, e_sum);
                        const double fac = (1.0 - psi) * (-c_power) / T;
#pragma unroll
                        for (int e = 0; e < cnt; ++e) scatter(awe_cost_layout[pos + 2 + e], fac * h * q * out[first + 1 + e]);
                        atomicAdd(gb + tf_v, fac * q * val / (double)pr.n_k);
                    } else {
                        b_sum = fma(q * val, val, b_sum);
#pragma unroll
                        for (int e = 0; e < cnt; ++e)
                            scatter(awe_cost_layout[pos + 2 + e], f_beta * 2.0 * q * val * out[first + 1 + e]);
                    }
                    pos += 2 + cnt;
                }
            }
        }
        __syncthreads();      // before the next intervals overwrite the staged blocks
    }
    // ---- weighted sums of squares per category: lanes -> warp -> L2 reductions
    double reg[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    auto add = [&](const int c, const double t) {
#pragma unroll
        for (int cc = 0; cc < 5; ++cc) reg[cc] += (c == cc) ? t : 0.0;
    };
#pragma unroll
    for (int ch = 0; ch < OBJ_NXC; ++ch) add(cx[ch], fx[ch] * sx[ch]), add(cd[ch], fd[ch] * sd[ch]);
#pragma unroll
    for (int ch = 0; ch < OBJ_NSC; ++ch) add(cs[ch], fs[ch] * ss[ch]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 5; ++c) reg[c] += __shfl_xor_sync(0xffffffffu, reg[c], o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        e_sum += __shfl_xor_sync(0xffffffffu, e_sum, o), b_sum += __shfl_xor_sync(0xffffffffu, b_sum, o);
    __shared__ double s_red[OBJ_LSQ_WARPS][OBJ_NACC];
    if (lane < 7) {
        double v = reg[0];
#pragma unroll
        for (int c = 1; c < 5; ++c) v = (lane == c) ? reg[c] : v;
        s_red[wid][lane] = (lane == 5) ? b_sum : (lane == 6) ? e_sum : v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    // ---- the instance's sums and the terms that do not belong to a node
    double ab[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        ab[c] = 0.0;
#pragma unroll
        for (int w = 0; w < OBJ_LSQ_WARPS; ++w) ab[c] += s_red[w][c];
    }
    double T_ref = 0.0;
    for (int i = 0; i < ob.n_tf; ++i) T_ref += ob.tf_coef[i] * __ldg(Pb + ob.tf_idx[i]);
    const double c_tf = __ldg(Pb + ob.p_cost + ob.c_tf);
    const double tracking = ab[0];
    const double beta_cost = f_beta * ab[5];
    const double time_cost = c_tf * (T - T_ref) * (T - T_ref);
    const double energy = ab[6];
    const double power_cost = -c_power * energy / T;
    double hom = 0.0;
    // the gradient entries below also receive reductions from the nodes: added the same way (no read of grad f)
    for (int n = 0; n < ob.n_phi_cost; ++n) {
        const double c = __ldg(Pb + ob.p_cost + ob.phi_cost[n]);
        hom += c * __ldg(Wb + pr.v_phi + n);
        atomicAdd(gb + pr.v_phi + n, c);
    }
    // general = fictitious + u_regularisation + xdot_regularisation + theta_regularisation + beta + time
    const double general = ab[3] + ab[2] + ab[1] + ab[4] + beta_cost + time_cost;
    F[b] = psi * tracking + (1.0 - psi) * power_cost + general + hom;
    for (int i = 0; i < ob.n_tf; ++i)
        atomicAdd(gb + ob.tf_idx[i], ob.tf_coef[i] * (2.0 * c_tf * (T - T_ref) - (1.0 - psi) * (-c_power) * energy / (T * T)));
    atomicAdd(gb + ob.psi_idx, tracking - power_cost);
    if (comp) {
        double *cb = comp + (size_t)b * 10;
        cb[0] = ab[0], cb[1] = ab[1], cb[2] = ab[2], cb[3] = ab[3], cb[4] = ab[4], cb[5] = beta_cost, cb[6] = time_cost;
        cb[7] = power_cost, cb[8] = hom, cb[9] = general;
    }
}

// --------------------------------------------------------------------------- tracking cost of the periodic MPC
// (awebox/pmpc.py:304-358; awebox_b200/ocp/objective.py: MpcObjective)  f = sum_i fac_i W_i (w_i - ref_i)^2 over the
// entries of V with a term (collocation states and algebraic variables, controls, terminal state), weights Q / R / P
// from Pmpc's parameter vector.  Element-wise in V: CTA = one instance, consecutive threads on consecutive entries
// (reads of w and ref, store of grad f coalesced), one sum per CTA.  Bound: HBM, 8 (3 n_v + 1) bytes per instance when
// every instance has its own reference, 8 (2 n_v + 1) with a shared one.
struct MpcObjParams {
    const int *kind, *idx;
    const double *fac;
    int p_ref, p_w[3];       // offsets of ref and of Q, R, P in the parameter vector
};

__global__ void __launch_bounds__(256)
awegpu_obj_mpc_kernel(const MpcObjParams mp, const int n_v, const double *__restrict__ W, const double *__restrict__ P,
                      const long long ldp, double *__restrict__ F, double *__restrict__ grad)
{
    const int b = blockIdx.x;
    const double *Wb = W + (size_t)b * n_v, *Pb = P + (size_t)b * ldp;
    double *gb = grad + (size_t)b * n_v;
    double sum = 0.0;
    for (int i = threadIdx.x; i < n_v; i += 256) {
        const int kd = __ldg(mp.kind + i);
        double g = 0.0;
        if (kd != 0) {
            const double wgt = (kd == 4 ? 1.0 : __ldg(Pb + mp.p_w[kd - 1] + __ldg(mp.idx + i))) * __ldg(mp.fac + i);
            const double diff = __ldg(Wb + i) - __ldg(Pb + mp.p_ref + i);
            sum = fma(wgt * diff, diff, sum);
            g = 2.0 * wgt * diff;
        }
        gb[i] = g;
    }
    __shared__ double s_sum[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_sum[w];
        F[b] = t;
    }
}

static constexpr size_t gather_smem_bytes() { return sizeof(double) * TILE * SLAB_STRIDE; }

// =========================================================================== host side / C ABI
struct awegpu_handle {
    int device = 0;
    DevProblem dp{};
    awegpu_problem_t host{};
    std::vector<void *> dev_allocs;
    std::vector<long long> colind, row;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    float last_ms = -1.f;
    long long launches = 0;
    std::string err;
    // staging for the host entry
    double *dW = nullptr, *dP = nullptr, *dG = nullptr, *dJ = nullptr;
    double *hW = nullptr, *hP = nullptr, *hG = nullptr, *hJ = nullptr;
    size_t capW = 0, capP = 0, capG = 0, capJ = 0, hcapW = 0, hcapP = 0, hcapG = 0, hcapJ = 0;
    cudaStream_t stream2 = nullptr, stream_in = nullptr, stream_out = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    // node-input records of one chunk of instances (see awegpu_gather_kernel)
    Scratch sc{};
    int chunk_tiles = 0, scratch_tiles = 0;
    int n_sm = 0;
    int lead_tiles = 2;           // tiles a program group may run ahead of the slowest group (AWEGPU_LEAD)
    int *counters = nullptr;      // [6][MAX_GROUPS] work counters of the persistent node kernels
    const int *colind32 = nullptr, *row32 = nullptr, *col32 = nullptr;     // CCS pattern on the device (awegpu_reduce_batch)
    double *dLam = nullptr, *dVec = nullptr, *dRed = nullptr;     // staging of awegpu_eval_batch_reduce_host
    size_t capLam = 0, capVec = 0, capRed = 0;
    std::vector<unsigned char> cta_group;     // [2 (full | residual only)][6][n_sm] group every CTA starts with
    int n_rest = 0;          // constant Jacobian entries the cooperative flush does not carry
    int n_affp = 0;          // parameter terms of the affine rows (> 0: the full parameter vector is required)
    // casadi sparsity buffers
    std::vector<long long> sp_v, sp_p, sp_g, sp_jv, sp_jp;
    std::vector<double> scratch_g;
    std::mutex casadi_mu;         // serialises the evaluations CasADi issues for this handle
    // objective path
    bool obj_enabled = false;
    ObjParams ob{};
    double *dF = nullptr, *dGrad = nullptr, *dComp = nullptr;
    size_t capF = 0, capGrad = 0, capComp = 0;
    const int *obj_rest = nullptr;      // entries of V outside the interval blocks
    int n_obj_rest = 0;
    bool mpc_obj = false;               // the objective is the tracking cost of the periodic MPC
    MpcObjParams mp{};
    int mpc_n_p = 0;
    // second-order hook (awegpu_hess.cu): tables uploaded by awegpu_hess_setup
    bool hess_ready = false, hess_obj = false;
    HessArgs hs{};
    std::vector<long long> hess_colind, hess_row, sp_hess, sp_lamf;
    std::vector<long long> sp_pf, sp_one, sp_jfv, sp_jfp, sp_out;     // CasADi sparsities of awegpu_f / jac_awegpu_f / awegpu_outputs
    std::vector<double> scratch_v, scratch_o;
    double *hess_scr = nullptr, *dHess = nullptr, *dLamH = nullptr;
    size_t cap_hess_scr = 0, capHess = 0, capLamH = 0;
    int hess_chunk_tiles = 64;
    float hess_ms = -1.f;
    cudaEvent_t ev_h0 = nullptr, ev_h1 = nullptr;
    // outputs map (awegpu_out.cu)
    double *dOutCoeff = nullptr, *dOut = nullptr;
    size_t capOut = 0;
    float out_ms = -1.f;
    cudaEvent_t ev_o0 = nullptr, ev_o1 = nullptr;
    // single-instance host calls as CUDA graphs (AWEGPU_GRAPH=0 disables)
    bool use_graph = true, capturing = false;
    cudaGraphExec_t graph_exec[8] = {nullptr};
};

static std::string g_create_error;
// CasADi external-function ABI: the handles that have been bound, addressed by the `mem` index CasADi passes to the
// evaluation functions (awegpu_g_checkout returns the slot of the handle bound at that time, so that two Functions
// of one library -- two OCPs of one model -- do not share state); evaluations of one handle are serialised
#define MAX_CASADI_MEM 64
static awegpu_handle *g_casadi_tab[MAX_CASADI_MEM] = {nullptr};
static int g_casadi_cur = -1, g_casadi_refs = 0;
static std::mutex g_casadi_mu;
static awegpu_handle *casadi_handle(int mem)
{
    if (mem < 0 || mem >= MAX_CASADI_MEM || g_casadi_tab[mem] == nullptr) mem = g_casadi_cur;
    return mem >= 0 ? g_casadi_tab[mem] : nullptr;
}

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
    {
        std::lock_guard<std::mutex> lock(g_casadi_mu);
        for (int i = 0; i < MAX_CASADI_MEM; ++i)
            if (g_casadi_tab[i] == h) {
                g_casadi_tab[i] = nullptr;
                if (g_casadi_cur == i) g_casadi_cur = -1;
            }
    }
    for (void *p : h->dev_allocs) cudaFree(p);
    cudaFree(h->dW), cudaFree(h->dP), cudaFree(h->dG), cudaFree(h->dJ);
    cudaFree(h->sc.rec), cudaFree(h->sc.xdc), cudaFree(h->sc.pt), cudaFree(h->sc.exc), cudaFree(h->sc.exs);
    cudaFree(h->counters), cudaFree(h->dLam), cudaFree(h->dVec), cudaFree(h->dRed);
    cudaFree(h->dF), cudaFree(h->dGrad), cudaFree(h->dComp);
    cudaFree(h->hess_scr), cudaFree(h->dHess), cudaFree(h->dLamH);
    cudaFree(h->dOutCoeff), cudaFree(h->dOut);
    if (h->ev_o0) cudaEventDestroy(h->ev_o0);
    if (h->ev_o1) cudaEventDestroy(h->ev_o1);
    if (h->ev_h0) cudaEventDestroy(h->ev_h0);
    if (h->ev_h1) cudaEventDestroy(h->ev_h1);
    for (cudaGraphExec_t ge : h->graph_exec)
        if (ge) cudaGraphExecDestroy(ge);
    cudaFreeHost(h->hW), cudaFreeHost(h->hP), cudaFreeHost(h->hG), cudaFreeHost(h->hJ);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_k0) cudaEventDestroy(h->ev_k0);
    if (h->ev_k1) cudaEventDestroy(h->ev_k1);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream_in) cudaStreamDestroy(h->stream_in);
    if (h->stream_out) cudaStreamDestroy(h->stream_out);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    delete h;
}

static int ensure_scratch(awegpu_handle *h, int tiles);
static int grow(awegpu_handle *h, double **dev, double **host, size_t *cap, size_t *hcap, size_t need, bool with_host);
static void drop_graphs(awegpu_handle *h);
static int grow_dev(awegpu_handle *h, double **dev, size_t *cap, size_t need);

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
        p->n_coll_out != AWE_N_COLL_OUT || p->n_shoot_out != AWE_N_SHOOT_OUT || p->n_kcoll_out != AWE_N_KCOLL_OUT ||
        p->n_kcoll_head != AWE_N_KCOLL_HEAD || p->kcoll_slot == nullptr ||
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
    d.n_coll_out = p->n_kcoll_out, d.n_shoot_out = p->n_shoot_out, d.n_shoot_eq = p->n_shoot_eq, d.n_ineq = p->n_ineq;
    d.n_periodic = p->n_periodic, d.row_periodic = p->row_periodic, d.x_last = p->x_last, d.x_first = p->x_base[0];
    d.cscale = p->constraint_scale;
    d.n_aff = p->n_aff;
    h->n_affp = p->n_aff ? p->affp_begin[p->n_aff] : 0;
    const int nk = p->n_k, dd = p->d;
    // destination codes of the node-local outputs: >= 0 slot of the CCS value array, -1 none,
    // <= -2 row (-2 - code) of g
    // (collocation nodes: the kernels' own output order, head outputs first, then the node's own CCS range)
    const int nco = p->n_kcoll_out;
    std::vector<int> coll_dest((size_t)nk * dd * nco), shoot_dest((size_t)nk * p->n_shoot_out);
    for (int kj = 0; kj < nk * dd; ++kj)
        for (int n = 0; n < nco; ++n)
            coll_dest[(size_t)kj * nco + n] = n < AWE_N_EQ ? -2 - (p->row_coll[kj] + n) : p->kcoll_slot[(size_t)kj * nco + n];
    // ---- own CCS ranges: chunk descriptors, positions of the outputs inside their chunk, constant entries
    std::vector<int4> own_desc((size_t)nk * dd * AWE_K_COLL_NOWNQ, make_int4(0, 0, 0, 0));
    std::vector<unsigned char> own_pos((size_t)nk * dd * (AWE_N_KCOLL_OWN > 0 ? AWE_N_KCOLL_OWN : 1), (unsigned char)OWN_DUMP);
    std::vector<uchar2> own_cst;
    std::vector<double> own_cval;
    std::vector<char> own_handled((size_t)p->nnz, 0);        // CCS slots written by the own-range flush
    {
        std::vector<int> const_of((size_t)p->nnz, -1);
        for (int e = 0; e < p->n_const; ++e) const_of[p->const_slot[e]] = e;
        std::vector<int> at;
        for (int kj = 0; kj < nk * dd && AWE_N_KCOLL_OWN > 0; ++kj) {
            const int k = kj / dd, j = kj % dd;
            const int vcol = p->x_base[k * (dd + 1) + j + 1];
            for (int q = 0; q < AWE_K_COLL_NOWNQ; ++q) {
                const long long s0 = p->colind[vcol + awe_k_coll_own_col0[q]];
                const long long s1 = p->colind[vcol + awe_k_coll_own_col0[q] + awe_k_coll_own_ncol[q]];
                const int len = (int)(s1 - s0);
                if (len > OWN_DUMP || len > 250) {
                    set_err(h, "own CCS range: a chunk does not fit the staging row");
                    return 4;
                }
                at.assign(len, -1);
                for (int n = awe_k_coll_own_first[q]; n < awe_k_coll_own_first[q] + awe_k_coll_own_cnt[q]; ++n) {
                    const int slot = p->kcoll_slot[(size_t)kj * nco + AWE_N_KCOLL_HEAD + n];
                    if (slot < 0) continue;
                    if (slot < s0 || slot >= s1 || at[slot - s0] != -1) {
                        set_err(h, "own CCS range: an output lies outside the columns of its chunk");
                        return 4;
                    }
                    at[slot - s0] = n;
                    own_pos[(size_t)kj * AWE_N_KCOLL_OWN + n] = (unsigned char)(slot - s0);
                }
                const int c0 = (int)own_cst.size();
                for (int e = 0; e < len; ++e) {
                    own_handled[s0 + e] = 1;
                    if (at[e] != -1) continue;
                    const int ce = const_of[s0 + e];
                    if (ce < 0) {
                        set_err(h, "own CCS range: a slot is neither an output nor a constant entry");
                        return 4;
                    }
                    int cid = -1;
                    for (size_t c = 0; c < own_cval.size(); ++c)
                        if (own_cval[c] == p->const_val[ce]) cid = (int)c;
                    if (cid < 0) own_cval.push_back(p->const_val[ce]), cid = (int)own_cval.size() - 1;
                    if (cid >= OWN_NCVAL || (int)own_cst.size() - c0 >= OWN_MAXCST) {
                        set_err(h, "own CCS range: too many constant entries");
                        return 4;
                    }
                    own_cst.push_back(make_uchar2((unsigned char)e, (unsigned char)cid));
                }
                own_desc[(size_t)kj * AWE_K_COLL_NOWNQ + q] = make_int4((int)s0, len, c0, (int)own_cst.size() - c0);
            }
        }
        own_cval.resize(OWN_NCVAL, 0.0);
    }
    for (int k = 0; k < nk; ++k)
        for (int n = 0; n < p->n_shoot_out; ++n)
            shoot_dest[(size_t)k * p->n_shoot_out + n] =
                n < p->n_shoot_eq                ? -2 - (p->row_shoot[k] + n)
                : n < p->n_shoot_eq + p->n_ineq ? -2 - (p->row_path[k] + n - p->n_shoot_eq)
                                                : p->shoot_slot[(size_t)k * p->n_shoot_out + n];
    // interval-cooperative flush tables (collocation kernel, stage 0).  Constant Jacobian entries (continuity
    // rows) that lie next to the staged entries of a chunk in the CCS value array ride along: their value comes
    // from a row of replicated constants, and the stores of the chunk become longer contiguous runs.  What does
    // not fit (periodicity, affine rows, x_{n_k}) stays with the gather kernel ("rest" lists).
    const int per = COOP_PER;
    std::vector<int2> coop_tab((size_t)nk * AWE_K_COLL_NCHUNKS * per, make_int2(-1, 0));
    std::vector<std::vector<std::pair<long long, int>>> ents((size_t)nk * AWE_K_COLL_NCHUNKS);
    std::vector<int> owner((size_t)p->nnz, -1);          // CCS slot -> (k, q) of the chunk that writes it
    for (int k = 0; k < nk; ++k)
        for (int q = 0; q < AWE_K_COLL_NCHUNKS; ++q) {
            std::vector<std::pair<long long, int>> &ent = ents[(size_t)k * AWE_K_COLL_NCHUNKS + q];
            for (int j = 0; j < dd; ++j)
                for (int o = 0; o < awe_k_coll_chunk_cnt[q]; ++o) {
                    const int code = coll_dest[((size_t)k * dd + j) * nco + awe_k_coll_chunk_first[q] + o];
                    if (code == -1) continue;
                    // g rows first (ascending row), then CCS slots (ascending)
                    const long long key = code <= -2 ? (long long)(-2 - code) : (1ll << 40) + code;
                    ent.push_back({key, j * AWE_CHUNK + o});
                    if (code >= 0) owner[code] = k * AWE_K_COLL_NCHUNKS + q;
                }
        }
    std::vector<double> cvals;
    std::vector<int> rest_slot, rest_begin;
    std::vector<double> rest_val;
    const bool coop_possible = (dd == NODE_WARPS);
    for (int k = 0; k < nk; ++k) {
        rest_begin.push_back((int)rest_slot.size());
        for (int e = p->const_begin[k]; e < p->const_begin[k + 1]; ++e) {
            const int slot = p->const_slot[e];
            const double v = p->const_val[e];
            if (own_handled[slot]) {
                // inside the own range of a collocation node: written by its flush; the gather kernel keeps it only
                // for the calls that do not run the full collocation programs
                continue;
            }
            int target = -1;
            for (int dist = 1; coop_possible && dist <= 2 && target < 0; ++dist)
                for (int sgn = -1; sgn <= 1 && target < 0; sgn += 2) {
                    const long long nb = (long long)slot + sgn * dist;
                    if (nb >= 0 && nb < p->nnz && owner[nb] >= 0 && ents[owner[nb]].size() < (size_t)per - 1) target = owner[nb];
                }
            int cid = -1;
            if (target >= 0) {
                for (size_t c = 0; c < cvals.size(); ++c)
                    if (cvals[c] == v) cid = (int)c;
                if (cid < 0 && cvals.size() < COOP_NCONST) cvals.push_back(v), cid = (int)cvals.size() - 1;
            }
            if (target >= 0 && cid >= 0)
                ents[target].push_back({(1ll << 40) + slot, COOP_STAGED + cid});
            else
                rest_slot.push_back(slot), rest_val.push_back(v);
        }
    }
    rest_begin.push_back((int)rest_slot.size());
    for (size_t t = 0; t < ents.size(); ++t) {
        std::sort(ents[t].begin(), ents[t].end());
        for (size_t e = 0; e < ents[t].size(); ++e) {
            const long long key = ents[t][e].first;
            coop_tab[t * per + e] = make_int2(key >= (1ll << 40) ? (int)(key - (1ll << 40)) : -2 - (int)key, ents[t][e].second);
        }
        coop_tab[t * per + per - 1] = make_int2(-1, (int)ents[t].size());      // never a real entry: at most per - 1
    }
    cvals.resize(COOP_NCONST, 0.0);
    h->n_rest = (int)rest_slot.size();
    int rc = 0;
    rc |= to_device(h, coop_tab.data(), coop_tab.size(), &d.coop_tab);
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
    rc |= to_device(h, rest_slot.data(), rest_slot.size(), &d.rest_slot);
    rc |= to_device(h, rest_begin.data(), rest_begin.size(), &d.rest_begin);
    rc |= to_device(h, rest_val.data(), rest_val.size(), &d.rest_val);
    rc |= to_device(h, cvals.data(), cvals.size(), &d.coop_const);
    rc |= to_device(h, own_desc.data(), own_desc.size(), &d.own_desc);
    rc |= to_device(h, own_pos.data(), own_pos.size(), &d.own_pos);
    rc |= to_device(h, own_cst.data(), own_cst.size(), &d.own_cst);
    rc |= to_device(h, own_cval.data(), own_cval.size(), &d.own_cval);
    rc |= to_device(h, p->aff_row, (size_t)p->n_aff, &d.aff_row);
    rc |= to_device(h, p->aff_begin, (size_t)p->n_aff + 1, &d.aff_begin);
    rc |= to_device(h, p->affp_begin, (size_t)p->n_aff + 1, &d.affp_begin);
    rc |= to_device(h, p->affp_col, (size_t)(p->n_aff ? p->affp_begin[p->n_aff] : 0), &d.affp_col);
    rc |= to_device(h, p->affp_coef, (size_t)(p->n_aff ? p->affp_begin[p->n_aff] : 0), &d.affp_coef);
    rc |= to_device(h, p->aff_col, (size_t)(p->n_aff ? p->aff_begin[p->n_aff] : 0), &d.aff_col);
    rc |= to_device(h, p->aff_coef, (size_t)(p->n_aff ? p->aff_begin[p->n_aff] : 0), &d.aff_coef);
    rc |= to_device(h, p->aff_const, (size_t)p->n_aff, &d.aff_const);
    if (rc) return rc;
    {
        // the collocation coefficients are constants of the library (one spec = one scheme and order): handles
        // that share the library must agree on them
        static std::mutex mu;
        static std::vector<double> seen;
        std::lock_guard<std::mutex> lock(mu);
        std::vector<double> now(p->coeff_coll, p->coeff_coll + (dd + 1) * (dd + 1));
        now.insert(now.end(), p->coeff_cont, p->coeff_cont + dd + 1);
        if (!seen.empty() && seen != now) {
            set_err(h, "another handle of this library uses different collocation coefficients");
            return 4;
        }
        seen = now;
    }
    CK(set_coeff_here(p->coeff_coll, p->coeff_cont));         // this unit's copy (gather kernel, residual-only kernels)
    h->colind.assign(p->colind, p->colind + p->n_v + 1);
    h->row.assign(p->row, p->row + p->nnz);
    {
        std::vector<int> c32(h->colind.begin(), h->colind.end()), r32(h->row.begin(), h->row.end()), col((size_t)p->nnz);
        for (int c = 0; c < p->n_v; ++c)
            for (long long e = p->colind[c]; e < p->colind[c + 1]; ++e) col[e] = c;
        if (to_device(h, c32.data(), c32.size(), &h->colind32) || to_device(h, r32.data(), r32.size(), &h->row32) ||
            to_device(h, col.data(), col.size(), &h->col32))
            return 2;
    }
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    CK(cudaEventCreate(&h->ev_k0));
    CK(cudaEventCreate(&h->ev_k1));
#define SET_SMEM(kern, bytes) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)))
    CK(awe_init_tu1(p->coeff_coll, p->coeff_cont));
    CK(awe_init_tu2(p->coeff_coll, p->coeff_cont));
    CK(awe_init_tu3(p->coeff_coll, p->coeff_cont));
    CK(awe_init_tu4(p->coeff_coll, p->coeff_cont));
    CK(cudaMalloc(&h->dOutCoeff, sizeof(double) * (N_D + 1) * (N_D + 1)));
    CK(cudaMemcpy(h->dOutCoeff, p->coeff_coll, sizeof(double) * (N_D + 1) * (N_D + 1), cudaMemcpyHostToDevice));
    CK(cudaEventCreate(&h->ev_o0));
    CK(cudaEventCreate(&h->ev_o1));
    CK((attr_node_impl<false, false, 0>()));
    CK((attr_node_impl<false, false, 1>()));
    CK((attr_node_impl<false, false, 2>()));
    CK((attr_node_impl<true, false, 0>()));
    CK((attr_node_impl<true, false, 1>()));
    CK((attr_node_impl<true, false, 2>()));
    SET_SMEM(awegpu_gather_kernel<false>, gather_smem_bytes());
    SET_SMEM(awegpu_obj_kernel, obj_smem_bytes(p->n_k));
    {
        // persistent node kernels: one CTA per SM; the group a CTA starts with, in proportion to the cost the
        // generator estimated for the groups of every (collocation | shooting, stage) launch
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        h->n_sm = prop.multiProcessorCount;
        if (getenv("AWEGPU_LEAD")) h->lead_tiles = atoi(getenv("AWEGPU_LEAD"));
        if (getenv("AWEGPU_GRAPH")) h->use_graph = atoi(getenv("AWEGPU_GRAPH")) != 0;
        const char *en = getenv("AWEGPU_CTAS");
        if (en && atoi(en) > 0) h->n_sm = atoi(en);
        static const int w_full[6][MAX_GROUPS] = {AWE_K_COLL_S0_WEIGHTS, AWE_K_COLL_S1_WEIGHTS, AWE_K_COLL_S2_WEIGHTS,
                                                  AWE_K_SHOOT_S0_WEIGHTS, AWE_K_SHOOT_S1_WEIGHTS, AWE_K_SHOOT_S2_WEIGHTS};
        static const int w_res[6][MAX_GROUPS] = {AWE_K_COLL_RES_S0_WEIGHTS, AWE_K_COLL_RES_S1_WEIGHTS, AWE_K_COLL_RES_S2_WEIGHTS,
                                                 AWE_K_SHOOT_RES_S0_WEIGHTS, AWE_K_SHOOT_RES_S1_WEIGHTS, AWE_K_SHOOT_RES_S2_WEIGHTS};
        const int n_full[6] = {AWE_K_COLL_S0_GROUPS, AWE_K_COLL_S1_GROUPS, AWE_K_COLL_S2_GROUPS,
                               AWE_K_SHOOT_S0_GROUPS, AWE_K_SHOOT_S1_GROUPS, AWE_K_SHOOT_S2_GROUPS};
        const int n_res[6] = {AWE_K_COLL_RES_S0_GROUPS, AWE_K_COLL_RES_S1_GROUPS, AWE_K_COLL_RES_S2_GROUPS,
                              AWE_K_SHOOT_RES_S0_GROUPS, AWE_K_SHOOT_RES_S1_GROUPS, AWE_K_SHOOT_RES_S2_GROUPS};
        if (h->n_sm > MAX_CTAS) h->n_sm = MAX_CTAS;
        std::vector<unsigned char> &cg = h->cta_group;
        cg.assign((size_t)2 * 6 * h->n_sm, 0);
        for (int f = 0; f < 2; ++f)
            for (int kind = 0; kind < 6; ++kind) {
                const int ng = f == 0 ? n_full[kind] : n_res[kind];
                const int *w = f == 0 ? w_full[kind] : w_res[kind];
                if (ng > MAX_GROUPS) {
                    set_err(h, "more program groups than MAX_GROUPS");
                    return 4;
                }
                double tot = 0, acc = 0;
                for (int g = 0; g < ng; ++g) tot += w[g] > 0 ? w[g] : 1;
                int g = 0;
                // CTA c starts with the group whose share of the cumulated weight contains (c + 0.5) / n_sm
                for (int c = 0; c < h->n_sm && ng > 0; ++c) {
                    const double x = (c + 0.5) / h->n_sm * tot;
                    while (g + 1 < ng && acc + (w[g] > 0 ? w[g] : 1) <= x) acc += (w[g] > 0 ? w[g] : 1), ++g;
                    cg[((size_t)f * 6 + kind) * h->n_sm + c] = (unsigned char)g;
                }
            }
        CK(cudaMalloc((void **)&h->counters, 6 * MAX_GROUPS * sizeof(int)));
    }
    SET_SMEM(awegpu_gather_kernel<AWE_BATCH_MINOR != 0>, gather_smem_bytes());
#undef SET_SMEM
    {
        // chunk of instances whose node-input records stay L2-resident while the program groups
        // stream over them (AWEGPU_CHUNK_TILES overrides; 16 tiles = 512 instances)
        const char *env = getenv("AWEGPU_CHUNK_TILES");
        int ct = env ? atoi(env) : 0;
        if (ct <= 0) {
            // default: the whole batch in one pass, capped by the memory the side buffers may take (2 GiB)
            const size_t per_tile = sizeof(double) * 32 * ((size_t)nk * (REC + (size_t)dd * (N_X + AWE_K_COLL_EXPORTS) +
                                                                         AWE_K_SHOOT_EXPORTS) + p->n_theta0);
            ct = (int)(((size_t)2 << 30) / per_tile);
            if (ct < 1) ct = 1;
        }
        h->chunk_tiles = ct;
    }
    if (p->obj_enabled == 2) {
        if (!p->mpc_kind || !p->mpc_idx || !p->mpc_fac || p->mpc_n_p <= 0) {
            set_err(h, "objective (mpc): the tables of the tracking cost are missing");
            return 4;
        }
        if (to_device(h, p->mpc_kind, (size_t)p->n_v, &h->mp.kind) || to_device(h, p->mpc_idx, (size_t)p->n_v, &h->mp.idx) ||
            to_device(h, p->mpc_fac, (size_t)p->n_v, &h->mp.fac))
            return 2;
        h->mp.p_ref = p->mpc_p_ref, h->mp.p_w[0] = p->mpc_p_Q, h->mp.p_w[1] = p->mpc_p_R, h->mp.p_w[2] = p->mpc_p_P;
        h->mpc_n_p = p->mpc_n_p;
        h->mpc_obj = true;
    } else if (p->obj_enabled) {
        ObjParams &ob = h->ob;
        if (p->n_phi > 16 || p->obj_n_tf > 2 || p->obj_n_tf < 1) {
            set_err(h, "objective: unsupported number of homotopy parameters / time-period terms");
            return 4;
        }
        if (to_device(h, p->obj_cat, (size_t)AWE_N_VAR, &ob.cat)) return 2;
        for (int c = 0; c < 6; ++c) ob.norm[c] = p->obj_norm[c];
        for (int j = 0; j < N_D; ++j) ob.quad[j] = p->obj_quad[j];
        ob.n_tf = p->obj_n_tf;
        for (int i = 0; i < p->obj_n_tf; ++i) ob.tf_idx[i] = p->obj_tf_idx[i], ob.tf_coef[i] = p->obj_tf_coef[i];
        ob.p_weights = p->obj_p_weights, ob.p_cost = p->obj_p_cost, ob.psi_idx = p->obj_psi_idx, ob.n_phi_cost = p->n_phi;
        for (int n = 0; n < p->n_phi; ++n) ob.phi_cost[n] = p->obj_phi_cost[n];
        ob.c_power = p->obj_c_power, ob.c_tf = p->obj_c_tf, ob.c_beta = p->obj_c_beta;
        for (int c = 0; c < 5; ++c) ob.c_cat[c] = p->obj_c_cat[c];
        {
            std::vector<char> in_block((size_t)p->n_v, 0);
            for (int k = 0; k < p->n_k; ++k)
                for (int c = 0; c < OFF_XNEXT; ++c) in_block[(size_t)p->x_base[k * (N_D + 1)] + c] = 1;
            std::vector<int> rest;
            for (int i = 0; i < p->n_v; ++i)
                if (!in_block[i]) rest.push_back(i);
            h->n_obj_rest = (int)rest.size();
            if (!rest.empty() && to_device(h, rest.data(), rest.size(), &h->obj_rest)) return 2;
        }
        h->obj_enabled = true;
    }
    // one-instance calls (the CasADi entry points, one IPOPT callback each) never allocate: side buffers of one tile
    // and the host-entry staging of one instance are created here; larger batches: awegpu_reserve
    if (ensure_scratch(h, 1)) return 2;
    if (grow(h, &h->dW, &h->hW, &h->capW, &h->hcapW, (size_t)p->n_v, true) | grow(h, &h->dP, &h->hP, &h->capP, &h->hcapP, (size_t)p->n_p, true) |
        grow(h, &h->dG, &h->hG, &h->capG, &h->hcapG, (size_t)p->n_g, true) | grow(h, &h->dJ, &h->hJ, &h->capJ, &h->hcapJ, (size_t)p->nnz, true))
        return 2;
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

// side buffers for `tiles` tiles (grown on demand, never shrunk; growth synchronises the device)
static int ensure_scratch(awegpu_handle *h, int tiles)
{
    if (tiles <= h->scratch_tiles) return 0;
    const DevProblem &d = h->dp;
    constexpr size_t nec = AWE_K_COLL_EXPORTS > AWE_K_COLL_RES_EXPORTS ? AWE_K_COLL_EXPORTS : AWE_K_COLL_RES_EXPORTS;
    constexpr size_t nes = AWE_K_SHOOT_EXPORTS > AWE_K_SHOOT_RES_EXPORTS ? AWE_K_SHOOT_EXPORTS : AWE_K_SHOOT_RES_EXPORTS;
    CK(cudaDeviceSynchronize());
    drop_graphs(h);
    cudaFree(h->sc.rec), cudaFree(h->sc.xdc), cudaFree(h->sc.pt), cudaFree(h->sc.exc), cudaFree(h->sc.exs);
    h->sc = Scratch{};
    h->scratch_tiles = 0;
    const size_t line = sizeof(double) * 32, t = (size_t)tiles, nk = (size_t)d.n_k;
    CK(cudaMalloc((void **)&h->sc.rec, line * t * nk * REC));
    CK(cudaMalloc((void **)&h->sc.xdc, line * t * nk * N_D * N_X));
    CK(cudaMalloc((void **)&h->sc.pt, line * t * (size_t)d.n_theta0));
    CK(cudaMalloc((void **)&h->sc.exc, line * t * nk * N_D * (nec ? nec : 1)));
    CK(cudaMalloc((void **)&h->sc.exs, line * t * nk * (nes ? nes : 1)));
    h->scratch_tiles = tiles;
    return 0;
}

static int grow(awegpu_handle *h, double **dev, double **host, size_t *cap, size_t *hcap, size_t need, bool with_host);

// side buffers (and, with AWEGPU_RESERVE_HOST, the staging buffers of the host entry) for batches of up to B
// instances, so that the evaluation calls that follow never allocate
extern "C" int awegpu_reserve(awegpu_handle_t *h, int32_t B, int32_t flags)
{
    if (!h || B <= 0) return 1;
    CK(cudaSetDevice(h->device));
    const int tiles = (B + TILE - 1) / TILE;
    int rc = ensure_scratch(h, tiles < h->chunk_tiles ? tiles : h->chunk_tiles);
    if (rc) return rc;
    if (flags & AWEGPU_RESERVE_HOST) {
        const DevProblem &d = h->dp;
        const bool stage = (flags & AWEGPU_HOST_PINNED) == 0;
        rc = grow(h, &h->dW, &h->hW, &h->capW, &h->hcapW, (size_t)B * d.n_v, stage) |
             grow(h, &h->dP, &h->hP, &h->capP, &h->hcapP, (size_t)B * d.n_p, stage) |
             grow(h, &h->dG, &h->hG, &h->capG, &h->hcapG, (size_t)B * d.n_g, stage) |
             grow(h, &h->dJ, &h->hJ, &h->capJ, &h->hcapJ, (size_t)B * d.nnz, stage);
    }
    return rc;
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
        if (h->n_affp > 0) {
            set_err(h, "this problem has rows that read p.ref (mpc initial conditions): pass full parameter vectors, not AWEGPU_P_THETA0");
            return 1;
        }
        ldp = (flags & AWEGPU_P_SHARED) ? 0 : h->dp.n_theta0;
        p_off = 0;
    } else {
        ldp = (flags & AWEGPU_P_SHARED) ? 0 : h->dp.n_p;
        p_off = h->dp.p_theta0;
    }
    double *jv = (flags & AWEGPU_SKIP_JAC) ? nullptr : Jv;
    const bool full = jv != nullptr, bm = (flags & AWEGPU_JAC_BATCH_MINOR) != 0;
    if (bm && !AWE_BATCH_MINOR) {
        set_err(h, "the batch-minor output layout is not compiled into the library of this (large) model");
        return 1;
    }
    const int tiles = (B + TILE - 1) / TILE;
    DevProblem dp = h->dp;
    if (full && !bm && N_D == NODE_WARPS) {
        // the collocation kernel of this call flushes cooperatively and carries most constant entries itself
        dp.const_slot = dp.rest_slot, dp.const_begin = dp.rest_begin, dp.const_val = dp.rest_val;
    }
    {
        const int rc_ = ensure_scratch(h, tiles < h->chunk_tiles ? tiles : h->chunk_tiles);
        if (rc_) return rc_;
    }
    const double *p_sh = (flags & AWEGPU_P_SHARED) ? P + p_off : nullptr;
    if (!h->capturing) CK(cudaEventRecord(h->ev0, st));
#define LAUNCH_STAGE(S_, F, T_, M, n_, st_)                                                                       \
    do {                                                                                                          \
        if (node_groups<S_, F, T_>() > 0) {                                                                       \
            NodeLaunch a_;                                                                                        \
            a_.n_ctas = h->n_sm, a_.st = st_, a_.dp = dp, a_.sc = h->sc, a_.B = B, a_.b_base = b_base;            \
            a_.n_items = n_, a_.p_sh = p_sh, a_.counters = h->counters;                                           \
            a_.cta_group = h->cta_group.data() + (size_t)((F) ? 0 : 6) * h->n_sm;                                        \
            a_.G = G, a_.Jv = jv;                                                                                 \
            a_.lead = h->lead_tiles * ((S_) ? dp.n_k : dp.n_k * N_D) / NODE_WARPS;                                \
            launch_node<S_, F, T_>(M, a_);                                                                        \
            h->launches += 1;                                                                                     \
        }                                                                                                         \
    } while (0)
#define LAUNCH_NODES(F, M)                                \
    do {                                                  \
        const int nc_ = ct * dp.n_k * N_D, ns_ = ct * dp.n_k; \
        /* one stream: running the shooting chain beside the collocation chain was measured and */ \
        /* bought nothing (every kernel fills the SMs on its own)                                */ \
        CK(cudaMemsetAsync(h->counters, 0, 6 * MAX_GROUPS * sizeof(int), st)); \
        LAUNCH_STAGE(false, F, 2, M, nc_, st);            \
        LAUNCH_STAGE(true, F, 2, M, ns_, st);             \
        LAUNCH_STAGE(false, F, 1, M, nc_, st);            \
        LAUNCH_STAGE(true, F, 1, M, ns_, st);             \
        LAUNCH_STAGE(true, F, 0, M, ns_, st);             \
        if (t0 == 0 && !h->capturing) CK(cudaEventRecord(h->ev_k0, st));   \
        LAUNCH_STAGE(false, F, 0, M, nc_, st);            \
        if (t0 == 0 && !h->capturing) CK(cudaEventRecord(h->ev_k1, st));   \
    } while (0)
    for (int t0 = 0; t0 < tiles; t0 += h->chunk_tiles) {
        const int ct = (tiles - t0 < h->chunk_tiles) ? (tiles - t0) : h->chunk_tiles;
        const int b_base = t0 * TILE;
        if (bm)
            awegpu_gather_kernel<AWE_BATCH_MINOR != 0><<<dim3(ct, dp.n_k), GATHER_THREADS, gather_smem_bytes(), st>>>(
                dp, h->sc, B, b_base, W, P, ldp, p_off, G, jv);
        else
            awegpu_gather_kernel<false><<<dim3(ct, dp.n_k), GATHER_THREADS, gather_smem_bytes(), st>>>(
                dp, h->sc, B, b_base, W, P, ldp, p_off, G, jv);
        if (full && !bm) LAUNCH_NODES(true, false);
        else if (full && bm) LAUNCH_NODES(true, true);
        else if (!full && !bm) LAUNCH_NODES(false, false);
        else LAUNCH_NODES(false, true);
        h->launches += 1;
    }
#undef LAUNCH_NODES
#undef LAUNCH_STAGE
    CK(cudaGetLastError());
    if (!h->capturing) {
        CK(cudaEventRecord(h->ev1, st));
        h->last_ms = -2.f;
    }
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

extern "C" float awegpu_last_kernel_ms(awegpu_handle_t *h)
{
    if (!h || h->last_ms == -1.f) return -1.f;
    if (cudaEventSynchronize(h->ev_k1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1) != cudaSuccess) return -1.f;
    return ms;
}

extern "C" int awegpu_kernel_info(int32_t *n_groups, int32_t *n_exports, int64_t *ops_coll, int64_t *ops_shoot)
{
    if (n_groups) {
        n_groups[0] = AWE_K_COLL_S0_GROUPS, n_groups[1] = AWE_K_COLL_S1_GROUPS, n_groups[2] = AWE_K_COLL_S2_GROUPS;
        n_groups[3] = AWE_K_SHOOT_S0_GROUPS, n_groups[4] = AWE_K_SHOOT_S1_GROUPS, n_groups[5] = AWE_K_SHOOT_S2_GROUPS;
    }
    if (n_exports) n_exports[0] = AWE_K_COLL_EXPORTS, n_exports[1] = AWE_K_SHOOT_EXPORTS;
    if (ops_coll) *ops_coll = AWE_K_COLL_OPS;
    if (ops_shoot) *ops_shoot = AWE_K_SHOOT_OPS;
    return 0;
}

// the CUDA graphs of the single-instance calls hold addresses of the staging and side buffers: dropped when one moves
static void drop_graphs(awegpu_handle *h)
{
    for (cudaGraphExec_t &ge : h->graph_exec)
        if (ge) cudaGraphExecDestroy(ge), ge = nullptr;
}

static int grow(awegpu_handle *h, double **dev, double **host, size_t *cap, size_t *hcap, size_t need, bool with_host)
{
    if (need > *cap || (with_host && need > *hcap)) drop_graphs(h);
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
        // caller buffers are page-locked: copy straight from / into them in chunks, H2D of chunk c+1,
        // the kernels of chunk c and D2H of chunk c-1 overlapping on three streams.  The kernels of all
        // chunks run on the handle's stream (they share the side buffers).
        if (flags & AWEGPU_JAC_BATCH_MINOR) {
            set_err(h, "AWEGPU_HOST_PINNED supports the instance-major layout only");
            return 1;
        }
        if (!h->stream_in) CK(cudaStreamCreateWithFlags(&h->stream_in, cudaStreamNonBlocking));
        if (!h->stream_out) CK(cudaStreamCreateWithFlags(&h->stream_out, cudaStreamNonBlocking));
        const bool shared = (flags & AWEGPU_P_SHARED) != 0;
        if (shared) CK(cudaMemcpyAsync(h->dP, P, nP * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        // instances per copy chunk (AWEGPU_COPY_CHUNK overrides): large enough for full-rate PCIe transfers, small
        // enough to keep the fill and drain of the three-stage pipeline short
        static const int chunk_env = getenv("AWEGPU_COPY_CHUNK") ? atoi(getenv("AWEGPU_COPY_CHUNK")) : 0;
        const int chunk = chunk_env > 0 ? chunk_env : 512;
        const int n_chunks = (B + chunk - 1) / chunk;
        while ((int)h->ev_pool.size() < 2 * n_chunks) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            h->ev_pool.push_back(e);
        }
        for (int c = 0; c < n_chunks; ++c) {
            const int b0 = c * chunk;
            const int nb = (B - b0 < chunk) ? (B - b0) : chunk;
            CK(cudaMemcpyAsync(h->dW + (size_t)b0 * d.n_v, W + (size_t)b0 * d.n_v, (size_t)nb * d.n_v * sizeof(double),
                               cudaMemcpyHostToDevice, h->stream_in));
            if (!shared)
                CK(cudaMemcpyAsync(h->dP + (size_t)b0 * p_len, P + (size_t)b0 * p_len, (size_t)nb * p_len * sizeof(double),
                                   cudaMemcpyHostToDevice, h->stream_in));
            CK(cudaEventRecord(h->ev_pool[2 * c], h->stream_in));
            CK(cudaStreamWaitEvent(h->stream, h->ev_pool[2 * c], 0));
            rc = awegpu_eval_batch(h, nb, h->dW + (size_t)b0 * d.n_v, shared ? h->dP : h->dP + (size_t)b0 * p_len,
                                   h->dG + (size_t)b0 * d.n_g, want_j ? h->dJ + (size_t)b0 * d.nnz : nullptr, flags, h->stream);
            if (rc) return rc;
            CK(cudaEventRecord(h->ev_pool[2 * c + 1], h->stream));
            CK(cudaStreamWaitEvent(h->stream_out, h->ev_pool[2 * c + 1], 0));
            CK(cudaMemcpyAsync(G + (size_t)b0 * d.n_g, h->dG + (size_t)b0 * d.n_g, (size_t)nb * d.n_g * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream_out));
            if (want_j)
                CK(cudaMemcpyAsync(Jv + (size_t)b0 * d.nnz, h->dJ + (size_t)b0 * d.nnz, (size_t)nb * d.nnz * sizeof(double),
                                   cudaMemcpyDeviceToHost, h->stream_out));
        }
        CK(cudaStreamSynchronize(h->stream_in));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaStreamSynchronize(h->stream_out));
        return 0;
    }
    memcpy(h->hW, W, nW * sizeof(double));
    memcpy(h->hP, P, nP * sizeof(double));
    if (B == 1 && h->use_graph) {
        // one instance per call (an IPOPT callback): the copies and the launches of the call are replayed as one CUDA
        // graph (all addresses are the handle's own staging buffers); one graph per flag combination
        const int key = (want_j ? 1 : 0) | ((flags & AWEGPU_P_THETA0) ? 2 : 0) | ((flags & AWEGPU_P_SHARED) ? 4 : 0);
        if (!h->graph_exec[key]) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            h->capturing = true;
            cudaError_t e1 = cudaMemcpyAsync(h->dW, h->hW, nW * sizeof(double), cudaMemcpyHostToDevice, h->stream);
            cudaError_t e2 = cudaMemcpyAsync(h->dP, h->hP, nP * sizeof(double), cudaMemcpyHostToDevice, h->stream);
            rc = awegpu_eval_batch(h, B, h->dW, h->dP, h->dG, want_j ? h->dJ : nullptr, flags, h->stream);
            cudaError_t e3 = cudaMemcpyAsync(h->hG, h->dG, nG * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
            cudaError_t e4 = want_j ? cudaMemcpyAsync(h->hJ, h->dJ, nJ * sizeof(double), cudaMemcpyDeviceToHost, h->stream) : cudaSuccess;
            h->capturing = false;
            cudaError_t e5 = cudaStreamEndCapture(h->stream, &graph);
            if (rc || e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                h->use_graph = false;                 // fall through to the plain path, for good
            } else {
                cudaError_t e6 = cudaGraphInstantiate(&h->graph_exec[key], graph, 0);
                cudaGraphDestroy(graph);
                if (e6 != cudaSuccess) h->graph_exec[key] = nullptr, h->use_graph = false, cudaGetLastError();
            }
        }
        if (h->graph_exec[key]) {
            CK(cudaGraphLaunch(h->graph_exec[key], h->stream));
            CK(cudaStreamSynchronize(h->stream));
            h->launches += 7;
            memcpy(G, h->hG, nG * sizeof(double));
            if (want_j) memcpy(Jv, h->hJ, nJ * sizeof(double));
            return 0;
        }
    }
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

// --------------------------------------------------------------------------- device-resident products (awegpu_reduce.cu)
extern "C" int awegpu_reduce_batch(awegpu_handle_t *h, int32_t B, const double *G, const double *Jvals, const double *lambda,
                                   const double *v, int32_t flags, double *JTl, double *Jv, double *cnorm, double *rnorm,
                                   double *ginf, void *stream)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    if (!Jvals || (ginf && !G)) {
        set_err(h, "awegpu_reduce_batch: null Jvals (or ginf without G)");
        return 1;
    }
    CK(cudaSetDevice(h->device));
    ReduceArgs a;
    a.B = B, a.n_v = h->dp.n_v, a.n_g = h->dp.n_g, a.nnz = h->dp.nnz;
    a.colind = h->colind32, a.row = h->row32, a.col = h->col32;
    a.G = G, a.J = Jvals, a.lambda = lambda, a.v = v;
    a.ld_lambda = (flags & AWEGPU_RED_LAMBDA_SHARED) ? 0 : h->dp.n_g;
    a.ld_v = (flags & AWEGPU_RED_V_SHARED) ? 0 : h->dp.n_v;
    a.JTl = JTl, a.Jv = Jv, a.cnorm = cnorm, a.rnorm = rnorm, a.ginf = ginf;
    const int rc = awe_reduce_launch(stream ? (cudaStream_t)stream : h->stream, a, h->n_sm);
    if (rc == -1) {
        set_err(h, "awegpu_reduce_batch: the row accumulators of this problem do not fit the shared memory of an SM");
        return 1;
    }
    if (rc != 0) {
        set_err(h, std::string("awegpu_reduce_batch: ") + cudaGetErrorString((cudaError_t)rc));
        return 2;
    }
    h->launches += 1;
    return 0;
}

static int grow_dev(awegpu_handle *h, double **dev, size_t *cap, size_t need)
{
    if (need > *cap) {
        if (*dev) cudaFree(*dev);
        *dev = nullptr, *cap = 0;
        CK(cudaMalloc((void **)dev, (need ? need : 1) * sizeof(double)));
        *cap = need;
    }
    return 0;
}

// host entry: g and jac_g are evaluated and kept on the device; only g and the requested products cross PCIe
extern "C" int awegpu_eval_batch_reduce_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P,
                                             const double *lambda, const double *v, int32_t flags, int32_t red_flags,
                                             double *G, double *JTl, double *Jv, double *cnorm, double *rnorm, double *ginf)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    if (flags & (AWEGPU_JAC_BATCH_MINOR | AWEGPU_SKIP_JAC)) {
        set_err(h, "awegpu_eval_batch_reduce_host: instance-major layout with Jacobian only");
        return 1;
    }
    CK(cudaSetDevice(h->device));
    const DevProblem &d = h->dp;
    const size_t nW = (size_t)B * d.n_v;
    const size_t p_len = (flags & AWEGPU_P_THETA0) ? d.n_theta0 : d.n_p;
    const size_t nP = (flags & AWEGPU_P_SHARED) ? p_len : (size_t)B * p_len;
    const size_t nG = (size_t)B * d.n_g, nJ = (size_t)B * d.nnz;
    const size_t nL = lambda ? ((red_flags & AWEGPU_RED_LAMBDA_SHARED) ? (size_t)d.n_g : nG) : 0;
    const size_t nV = v ? ((red_flags & AWEGPU_RED_V_SHARED) ? (size_t)d.n_v : nW) : 0;
    // results: [JTl B n_v | Jv B n_g | cnorm B n_v | rnorm B n_g | ginf B]
    const size_t oJTl = 0, oJv = oJTl + (JTl ? nW : 0), oCn = oJv + (Jv ? nG : 0), oRn = oCn + (cnorm ? nW : 0),
                 oGi = oRn + (rnorm ? nG : 0), nRed = oGi + (ginf ? (size_t)B : 0);
    // instances per pipeline chunk: H2D of chunk c+1, the kernels of chunk c and D2H of chunk c-1 overlap on three
    // streams (caller buffers should be page-locked for that; pageable buffers work, without overlap)
    static const int chunk_env = getenv("AWEGPU_COPY_CHUNK") ? atoi(getenv("AWEGPU_COPY_CHUNK")) : 0;
    const int chunk = chunk_env > 0 ? chunk_env : 1024;         // measured: 256 / 512 / 1024 / 2048 -> 21.6 / 15.0 / 12.6 / 13.9 ms (AP2, B=4096)
    const int n_chunks = (B + chunk - 1) / chunk;
    const size_t cb = (size_t)(B < chunk ? B : chunk);
    int rc = grow(h, &h->dW, &h->hW, &h->capW, &h->hcapW, nW, false) | grow(h, &h->dP, &h->hP, &h->capP, &h->hcapP, nP, false) |
             grow(h, &h->dG, &h->hG, &h->capG, &h->hcapG, nG, false) |
             grow(h, &h->dJ, &h->hJ, &h->capJ, &h->hcapJ, cb * d.nnz, false) |       // one chunk: consumed on the device
             grow_dev(h, &h->dLam, &h->capLam, nL) | grow_dev(h, &h->dVec, &h->capVec, nV) | grow_dev(h, &h->dRed, &h->capRed, nRed);
    if (rc) return rc;
    if (!h->stream_in) CK(cudaStreamCreateWithFlags(&h->stream_in, cudaStreamNonBlocking));
    if (!h->stream_out) CK(cudaStreamCreateWithFlags(&h->stream_out, cudaStreamNonBlocking));
    while ((int)h->ev_pool.size() < 2 * n_chunks) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_pool.push_back(e);
    }
    const bool p_shared = (flags & AWEGPU_P_SHARED) != 0, l_shared = (red_flags & AWEGPU_RED_LAMBDA_SHARED) != 0,
               v_shared = (red_flags & AWEGPU_RED_V_SHARED) != 0;
    cudaStream_t st = h->stream;
    if (p_shared) CK(cudaMemcpyAsync(h->dP, P, nP * sizeof(double), cudaMemcpyHostToDevice, st));
    if (nL && l_shared) CK(cudaMemcpyAsync(h->dLam, lambda, nL * sizeof(double), cudaMemcpyHostToDevice, st));
    if (nV && v_shared) CK(cudaMemcpyAsync(h->dVec, v, nV * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int c = 0; c < n_chunks; ++c) {
        const size_t b0 = (size_t)c * chunk;
        const size_t nb = (size_t)B - b0 < (size_t)chunk ? (size_t)B - b0 : (size_t)chunk;
        CK(cudaMemcpyAsync(h->dW + b0 * d.n_v, W + b0 * d.n_v, nb * d.n_v * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        if (!p_shared)
            CK(cudaMemcpyAsync(h->dP + b0 * p_len, P + b0 * p_len, nb * p_len * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        if (nL && !l_shared)
            CK(cudaMemcpyAsync(h->dLam + b0 * d.n_g, lambda + b0 * d.n_g, nb * d.n_g * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        if (nV && !v_shared)
            CK(cudaMemcpyAsync(h->dVec + b0 * d.n_v, v + b0 * d.n_v, nb * d.n_v * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CK(cudaEventRecord(h->ev_pool[2 * c], h->stream_in));
        CK(cudaStreamWaitEvent(st, h->ev_pool[2 * c], 0));
        rc = awegpu_eval_batch(h, (int)nb, h->dW + b0 * d.n_v, p_shared ? h->dP : h->dP + b0 * p_len, h->dG + b0 * d.n_g, h->dJ, flags, st);
        if (rc) return rc;
        rc = awegpu_reduce_batch(h, (int)nb, h->dG + b0 * d.n_g, h->dJ, nL ? (l_shared ? h->dLam : h->dLam + b0 * d.n_g) : nullptr,
                                 nV ? (v_shared ? h->dVec : h->dVec + b0 * d.n_v) : nullptr, red_flags,
                                 JTl ? h->dRed + oJTl + b0 * d.n_v : nullptr, Jv ? h->dRed + oJv + b0 * d.n_g : nullptr,
                                 cnorm ? h->dRed + oCn + b0 * d.n_v : nullptr, rnorm ? h->dRed + oRn + b0 * d.n_g : nullptr,
                                 ginf ? h->dRed + oGi + b0 : nullptr, st);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev_pool[2 * c + 1], st));
        CK(cudaStreamWaitEvent(h->stream_out, h->ev_pool[2 * c + 1], 0));
        cudaStream_t so = h->stream_out;
        if (G) CK(cudaMemcpyAsync(G + b0 * d.n_g, h->dG + b0 * d.n_g, nb * d.n_g * sizeof(double), cudaMemcpyDeviceToHost, so));
        if (JTl) CK(cudaMemcpyAsync(JTl + b0 * d.n_v, h->dRed + oJTl + b0 * d.n_v, nb * d.n_v * sizeof(double), cudaMemcpyDeviceToHost, so));
        if (Jv) CK(cudaMemcpyAsync(Jv + b0 * d.n_g, h->dRed + oJv + b0 * d.n_g, nb * d.n_g * sizeof(double), cudaMemcpyDeviceToHost, so));
        if (cnorm) CK(cudaMemcpyAsync(cnorm + b0 * d.n_v, h->dRed + oCn + b0 * d.n_v, nb * d.n_v * sizeof(double), cudaMemcpyDeviceToHost, so));
        if (rnorm) CK(cudaMemcpyAsync(rnorm + b0 * d.n_g, h->dRed + oRn + b0 * d.n_g, nb * d.n_g * sizeof(double), cudaMemcpyDeviceToHost, so));
        if (ginf) CK(cudaMemcpyAsync(ginf + b0, h->dRed + oGi + b0, nb * sizeof(double), cudaMemcpyDeviceToHost, so));
    }
    CK(cudaStreamSynchronize(h->stream_in));
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamSynchronize(h->stream_out));
    return 0;
}

// --------------------------------------------------------------------------- objective
extern "C" int awegpu_eval_objective(awegpu_handle_t *h, int32_t B, const double *W, const double *P, int32_t flags, double *F,
                                     double *GradF, double *comp, void *stream)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    if (h->mpc_obj) {
        if (!W || !P || !F || !GradF || comp || (flags & AWEGPU_P_THETA0)) {
            set_err(h, "awegpu_eval_objective (mpc): W, P (Pmpc's parameter vectors), F and GradF are required, comp must be NULL");
            return 1;
        }
        CK(cudaSetDevice(h->device));
        cudaStream_t st_ = stream ? (cudaStream_t)stream : h->stream;
        awegpu_obj_mpc_kernel<<<B, 256, 0, st_>>>(h->mp, h->dp.n_v, W, P, (flags & AWEGPU_P_SHARED) ? 0 : h->mpc_n_p, F, GradF);
        h->launches += 1;
        CK(cudaGetLastError());
        return 0;
    }
    if (!h->obj_enabled) {
        set_err(h, "awegpu_eval_objective: this problem has no objective on this path");
        return 1;
    }
    if (!W || !P || !F || !GradF || (flags & AWEGPU_P_THETA0)) {
        set_err(h, "awegpu_eval_objective: W, P (full parameter vectors), F and GradF are required");
        return 1;
    }
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const DevProblem &dp = h->dp;
    const long long ldp = (flags & AWEGPU_P_SHARED) ? 0 : dp.n_p;
    // one launch: every entry of GradF is written (interval blocks) or zeroed (the rest) by the kernel itself
    awegpu_obj_kernel<<<B, 32 * OBJ_LSQ_WARPS, obj_smem_bytes(dp.n_k), st>>>(dp, h->ob, B, W, P, ldp, dp.p_theta0,
                                                                             obj_stage_k(dp.n_k), h->obj_rest, h->n_obj_rest,
                                                                             F, GradF, comp);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int awegpu_eval_objective_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P, int32_t flags,
                                          double *F, double *GradF, double *comp)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    CK(cudaSetDevice(h->device));
    const DevProblem &d = h->dp;
    const size_t n_p = h->mpc_obj ? (size_t)h->mpc_n_p : (size_t)d.n_p;
    const size_t nW = (size_t)B * d.n_v, nP = (flags & AWEGPU_P_SHARED) ? n_p : (size_t)B * n_p;
    int rc = grow(h, &h->dW, &h->hW, &h->capW, &h->hcapW, nW, false) | grow(h, &h->dP, &h->hP, &h->capP, &h->hcapP, nP, false) |
             grow_dev(h, &h->dF, &h->capF, (size_t)B) | grow_dev(h, &h->dGrad, &h->capGrad, nW) |
             (comp ? grow_dev(h, &h->dComp, &h->capComp, (size_t)B * 10) : 0);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(h->dW, W, nW * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->dP, P, nP * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = awegpu_eval_objective(h, B, h->dW, h->dP, flags, h->dF, h->dGrad, comp ? h->dComp : nullptr, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(F, h->dF, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(GradF, h->dGrad, nW * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (comp) CK(cudaMemcpyAsync(comp, h->dComp, (size_t)B * 10 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}


// --------------------------------------------------------------------------- second-order hook
extern "C" int awegpu_hess_setup(awegpu_handle_t *h, int64_t nnz, int32_t n_hv, int64_t n_asm, const int64_t *colind,
                                 const int64_t *row, const int64_t *asm_ptr, const int64_t *asm_idx, const double *asm_coef,
                                 int32_t v_x0, int32_t row0, int32_t row_stride, const double *coeff, const double *quad,
                                 int32_t n_ec, int32_t n_es, int32_t with_objective, int32_t n_t, const int32_t *t_idx,
                                 const double *t_coef)
{
    if (!h || !colind || !row || !asm_ptr || !asm_idx || !asm_coef || !coeff || !quad || nnz <= 0 || n_asm <= 0) return 1;
    {
        int ec = 0, es = 0;
        awe_hess_node_dims(&ec, &es);
        if (ec != n_ec || es != n_es || (with_objective != 0) != (awe_hess_has_objective() != 0)) {
            set_err(h, "awegpu_hess_setup: the tables do not belong to the node Hessian programs of this library");
            return 1;
        }
        if (n_t < 0 || n_t > 2 || (with_objective && (n_t == 0 || !t_idx || !t_coef))) {
            set_err(h, "awegpu_hess_setup: time period terms");
            return 1;
        }
    }
    if (n_asm >= (1LL << 31) || nnz >= (1LL << 31)) {
        set_err(h, "awegpu_hess_setup: table too large for 32-bit indices");
        return 1;
    }
    CK(cudaSetDevice(h->device));
    const DevProblem &d = h->dp;
    h->hess_colind.assign(colind, colind + d.n_v + 1);
    h->hess_row.assign(row, row + nnz);
    std::vector<int> ptr32(asm_ptr, asm_ptr + nnz + 1), idx32(asm_idx, asm_idx + n_asm);
    HessArgs &a = h->hs;
    a = HessArgs{};
    int rc = to_device(h, ptr32.data(), ptr32.size(), &a.asm_ptr) | to_device(h, idx32.data(), idx32.size(), &a.asm_idx) |
             to_device(h, asm_coef, (size_t)n_asm, &a.asm_coef);
    if (rc) return rc;
    a.n_v = d.n_v, a.n_g = d.n_g, a.n_k = d.n_k, a.n_hv = n_hv, a.nnz = nnz;
    a.v_x0 = v_x0, a.v_phi = d.v_phi, a.row0 = row0, a.row_stride = row_stride;
    a.theta_idx = d.theta_idx;
    a.n_var = h->host.n_var, a.n_t = with_objective ? n_t : 0;
    for (int i = 0; i < a.n_t; ++i) a.t_idx[i] = t_idx[i], a.t_coef[i] = t_coef[i];
    {
        std::vector<double> z((size_t)d.n_p + 1, 0.0);
        rc = to_device(h, z.data(), z.size(), &a.zeros);
        if (rc) return rc;
    }
    h->hess_obj = with_objective != 0;
    if (awe_hess_set_coeff(coeff, quad)) {
        set_err(h, "awegpu_hess_setup: cudaMemcpyToSymbol failed");
        return 2;
    }
    h->sp_hess.clear();
    h->sp_hess.push_back(d.n_v);
    h->sp_hess.push_back(d.n_v);
    h->sp_hess.insert(h->sp_hess.end(), h->hess_colind.begin(), h->hess_colind.end());
    h->sp_hess.insert(h->sp_hess.end(), h->hess_row.begin(), h->hess_row.end());
    h->sp_lamf = {1, 1, 1};
    if (!h->ev_h0) CK(cudaEventCreate(&h->ev_h0));
    if (!h->ev_h1) CK(cudaEventCreate(&h->ev_h1));
    h->hess_ready = true;
    return 0;
}

extern "C" int awegpu_hess_has_objective(void) { return awe_hess_has_objective(); }

extern "C" int64_t awegpu_hess_nnz(awegpu_handle_t *h) { return (h && h->hess_ready) ? h->hs.nnz : -1; }

extern "C" int awegpu_hess_sparsity(awegpu_handle_t *h, int64_t *colind, int64_t *row)
{
    if (!h || !h->hess_ready) return 1;
    for (size_t i = 0; i < h->hess_colind.size(); ++i) colind[i] = h->hess_colind[i];
    for (size_t i = 0; i < h->hess_row.size(); ++i) row[i] = h->hess_row[i];
    return 0;
}

extern "C" int awegpu_eval_hessian(awegpu_handle_t *h, int32_t B, const double *W, const double *P, const double *Lam,
                                   const double *Sigma, int32_t flags, double *H, void *stream)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    if (!h->hess_ready) {
        set_err(h, "awegpu_eval_hessian: awegpu_hess_setup has not been called for this handle");
        return 1;
    }
    if (!W || !P || !Lam || !H) {
        set_err(h, "awegpu_eval_hessian: W, P, Lam and H are required");
        return 1;
    }
    if (Sigma && !h->hess_obj) {
        set_err(h, "awegpu_eval_hessian: this problem has no objective part on this path (pass Sigma = NULL)");
        return 1;
    }
    if (Sigma && (flags & AWEGPU_P_THETA0)) {
        set_err(h, "awegpu_eval_hessian: the objective part needs full parameter vectors (not AWEGPU_P_THETA0)");
        return 1;
    }
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const DevProblem &d = h->dp;
    const int tiles = (B + TILE - 1) / TILE;
    const int chunk = tiles < h->hess_chunk_tiles ? tiles : h->hess_chunk_tiles;
    const size_t need = awe_hess_scratch_per_tile(h->hs.n_hv) / sizeof(double) * (size_t)chunk;
    if (need > h->cap_hess_scr) {
        CK(cudaStreamSynchronize(st));
        int rc = grow_dev(h, &h->hess_scr, &h->cap_hess_scr, need);
        if (rc) return rc;
    }
    HessArgs a = h->hs;
    a.B = B, a.W = W, a.P = P, a.Lam = Lam, a.Sigma = Sigma, a.H = H, a.scr = h->hess_scr;
    a.ld_p = (flags & AWEGPU_P_SHARED) ? 0 : ((flags & AWEGPU_P_THETA0) ? d.n_theta0 : d.n_p);
    a.p_off = (flags & AWEGPU_P_THETA0) ? 0 : d.p_theta0;
    CK(cudaEventRecord(h->ev_h0, st));
    for (int t0 = 0; t0 < tiles; t0 += chunk) {
        const int nt = tiles - t0 < chunk ? tiles - t0 : chunk;
        const int e = awe_hess_launch(st, a, t0, nt);
        if (e) {
            set_err(h, std::string("awegpu_eval_hessian: ") + cudaGetErrorString((cudaError_t)e));
            return 2;
        }
        h->launches += 2;
    }
    CK(cudaEventRecord(h->ev_h1, st));
    h->hess_ms = -2.f;      // recorded, not read yet
    return 0;
}

extern "C" float awegpu_last_hessian_ms(awegpu_handle_t *h)
{
    if (!h || h->hess_ms == -1.f) return -1.f;
    if (cudaEventSynchronize(h->ev_h1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, h->ev_h0, h->ev_h1) != cudaSuccess) return -1.f;
    return ms;
}

extern "C" int awegpu_eval_hessian_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P, const double *Lam,
                                        const double *Sigma, int32_t flags, double *H)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    if (!h->hess_ready) {
        set_err(h, "awegpu_eval_hessian_host: awegpu_hess_setup has not been called for this handle");
        return 1;
    }
    CK(cudaSetDevice(h->device));
    const DevProblem &d = h->dp;
    const size_t ldp = (flags & AWEGPU_P_THETA0) ? (size_t)d.n_theta0 : (size_t)d.n_p;
    const size_t nW = (size_t)B * d.n_v, nP = (flags & AWEGPU_P_SHARED) ? ldp : (size_t)B * ldp;
    const size_t nL = (size_t)B * d.n_g, nH = (size_t)B * (size_t)h->hs.nnz;
    int rc = grow(h, &h->dW, &h->hW, &h->capW, &h->hcapW, nW, false) | grow(h, &h->dP, &h->hP, &h->capP, &h->hcapP, nP, false) |
             grow_dev(h, &h->dLamH, &h->capLamH, nL + (size_t)B) | grow_dev(h, &h->dHess, &h->capHess, nH);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(h->dW, W, nW * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->dP, P, nP * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->dLamH, Lam, nL * sizeof(double), cudaMemcpyHostToDevice, st));
    if (Sigma) CK(cudaMemcpyAsync(h->dLamH + nL, Sigma, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = awegpu_eval_hessian(h, B, h->dW, h->dP, h->dLamH, Sigma ? h->dLamH + nL : nullptr, flags, h->dHess, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(H, h->dHess, nH * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

// --------------------------------------------------------------------------- outputs map (awegpu_out.cu)
extern "C" int32_t awegpu_outputs_n(void) { return awe_out_n(); }
extern "C" const char *awegpu_outputs_label(int32_t i) { return awe_out_label(i); }
extern "C" int32_t awegpu_outputs_nodes(awegpu_handle_t *h) { return h ? h->dp.n_k * (N_D + 1) : -1; }

extern "C" int awegpu_eval_outputs(awegpu_handle_t *h, int32_t B, const double *W, const double *P, int32_t flags, double *Out,
                                   void *stream)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    if (!W || !P || !Out) {
        set_err(h, "awegpu_eval_outputs: W, P and Out are required");
        return 1;
    }
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const DevProblem &d = h->dp;
    OutArgs a{};
    a.B = B, a.n_v = d.n_v, a.n_k = d.n_k, a.W = W, a.P = P, a.Out = Out;
    a.ld_p = (flags & AWEGPU_P_SHARED) ? 0 : ((flags & AWEGPU_P_THETA0) ? d.n_theta0 : d.n_p);
    a.p_off = (flags & AWEGPU_P_THETA0) ? 0 : d.p_theta0;
    a.v_x0 = d.x_first, a.v_phi = d.v_phi, a.theta_idx = d.theta_idx, a.tf_idx = d.tf_idx, a.coeff = h->dOutCoeff;
    CK(cudaEventRecord(h->ev_o0, st));
    const int e = awe_out_launch(st, a);
    if (e) {
        set_err(h, std::string("awegpu_eval_outputs: ") + cudaGetErrorString((cudaError_t)e));
        return 2;
    }
    h->launches += 1;
    CK(cudaEventRecord(h->ev_o1, st));
    h->out_ms = -2.f;
    return 0;
}

extern "C" float awegpu_last_outputs_ms(awegpu_handle_t *h)
{
    if (!h || h->out_ms == -1.f) return -1.f;
    if (cudaEventSynchronize(h->ev_o1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, h->ev_o0, h->ev_o1) != cudaSuccess) return -1.f;
    return ms;
}

extern "C" int awegpu_eval_outputs_host(awegpu_handle_t *h, int32_t B, const double *W, const double *P, int32_t flags,
                                        double *Out)
{
    if (!h) return 1;
    if (B <= 0) return 0;
    CK(cudaSetDevice(h->device));
    const DevProblem &d = h->dp;
    const size_t ldp = (flags & AWEGPU_P_THETA0) ? (size_t)d.n_theta0 : (size_t)d.n_p;
    const size_t nW = (size_t)B * d.n_v, nP = (flags & AWEGPU_P_SHARED) ? ldp : (size_t)B * ldp;
    const size_t nO = (size_t)B * (size_t)awe_out_n() * (size_t)(d.n_k * (N_D + 1));
    int rc = grow(h, &h->dW, &h->hW, &h->capW, &h->hcapW, nW, false) | grow(h, &h->dP, &h->hP, &h->capP, &h->hcapP, nP, false) |
             grow_dev(h, &h->dOut, &h->capOut, nO);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(h->dW, W, nW * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->dP, P, nP * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = awegpu_eval_outputs(h, B, h->dW, h->dP, flags, h->dOut, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(Out, h->dOut, nO * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
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
    std::lock_guard<std::mutex> lock(g_casadi_mu);
    int slot = -1;
    for (int i = 0; i < MAX_CASADI_MEM && slot < 0; ++i)
        if (g_casadi_tab[i] == h) slot = i;
    for (int i = 0; i < MAX_CASADI_MEM && slot < 0; ++i)
        if (g_casadi_tab[i] == nullptr) slot = i;
    if (slot < 0) {
        set_err(h, "awegpu_casadi_bind: no free slot");
        return 1;
    }
    g_casadi_tab[slot] = h, g_casadi_cur = slot;
    dense_sp(h->sp_v, h->dp.n_v);
    dense_sp(h->sp_p, h->dp.n_p);
    dense_sp(h->sp_g, h->dp.n_g);
    h->sp_jv.clear();
    h->sp_jv.push_back(h->dp.n_g);
    h->sp_jv.push_back(h->dp.n_v);
    h->sp_jv.insert(h->sp_jv.end(), h->colind.begin(), h->colind.end());
    h->sp_jv.insert(h->sp_jv.end(), h->row.begin(), h->row.end());
    // jac_g_P is declared structurally empty: nlpsol only differentiates g with respect to V.  g does depend on P
    // (theta0 everywhere, p.ref in the initial-condition rows of an mpc trial): this library is NOT a source of
    // parametric sensitivities.
    h->sp_jp.assign(2 + (size_t)h->dp.n_p + 1, 0);      // n_g x n_p, no non-zeros
    h->sp_jp[0] = h->dp.n_g;
    h->sp_jp[1] = h->dp.n_p;
    return 0;
}

static int casadi_eval(const double **arg, double **res, bool jac, int mem)
{
    awegpu_handle *h = casadi_handle(mem);
    if (!h) return 1;
    if (!arg || !arg[0] || !arg[1]) {
        set_err(h, "awegpu_g: V and P must be given (NULL means all-zero, for which g is undefined)");
        return 1;
    }
    double *out = res ? res[0] : nullptr;
    if (!out) return 0;
    std::lock_guard<std::mutex> lock(h->casadi_mu);
    if (jac) h->scratch_g.resize(h->dp.n_g);
    return awegpu_eval_batch_host(h, 1, arg[0], arg[1], jac ? h->scratch_g.data() : out, jac ? out : nullptr,
                                  jac ? 0 : AWEGPU_SKIP_JAC);
}

// optional memory-management entry points of a CasADi external (casadi/core/external.cpp looks them up by name):
// checkout -> index of the handle that is bound now; release: nothing to return; incref / decref: bookkeeping only
// (handles are owned by whoever called awegpu_create); alloc_mem / init_mem / free_mem: the same slots
#define CASADI_MEM_FUNCS(F)                                                                                  \
    extern "C" int F##_checkout(void) { std::lock_guard<std::mutex> lock(g_casadi_mu); return g_casadi_cur < 0 ? 0 : g_casadi_cur; } \
    extern "C" void F##_release(int) {}                                                                      \
    extern "C" void F##_incref(void) { std::lock_guard<std::mutex> lock(g_casadi_mu); ++g_casadi_refs; }     \
    extern "C" void F##_decref(void) { std::lock_guard<std::mutex> lock(g_casadi_mu); --g_casadi_refs; }     \
    extern "C" int F##_alloc_mem(void) { return F##_checkout(); }                                            \
    extern "C" int F##_init_mem(int mem) { return casadi_handle(mem) ? 0 : 1; }                              \
    extern "C" void F##_free_mem(int) {}
CASADI_MEM_FUNCS(awegpu_g)
CASADI_MEM_FUNCS(jac_awegpu_g)

extern "C" int awegpu_g(const double **arg, double **res, long long *, double *, int mem) { return casadi_eval(arg, res, false, mem); }
extern "C" long long awegpu_g_n_in(void) { return 2; }
extern "C" long long awegpu_g_n_out(void) { return 1; }
extern "C" const char *awegpu_g_name_in(long long i) { return i == 0 ? "V" : i == 1 ? "P" : nullptr; }
extern "C" const char *awegpu_g_name_out(long long i) { return i == 0 ? "g" : nullptr; }
extern "C" const long long *awegpu_g_sparsity_in(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    if (!h) return nullptr;
    return i == 0 ? h->sp_v.data() : i == 1 ? h->sp_p.data() : nullptr;
}
extern "C" const long long *awegpu_g_sparsity_out(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    return (h && i == 0) ? h->sp_g.data() : nullptr;
}
extern "C" int awegpu_g_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w)
{
    if (sz_arg) *sz_arg = 2;
    if (sz_res) *sz_res = 1;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}

extern "C" int jac_awegpu_g(const double **arg, double **res, long long *, double *, int mem) { return casadi_eval(arg, res, true, mem); }
extern "C" long long jac_awegpu_g_n_in(void) { return 3; }
extern "C" long long jac_awegpu_g_n_out(void) { return 2; }
extern "C" const char *jac_awegpu_g_name_in(long long i) { return i == 0 ? "V" : i == 1 ? "P" : i == 2 ? "out_g" : nullptr; }
extern "C" const char *jac_awegpu_g_name_out(long long i) { return i == 0 ? "jac_g_V" : i == 1 ? "jac_g_P" : nullptr; }
extern "C" const long long *jac_awegpu_g_sparsity_in(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    if (!h) return nullptr;
    return i == 0 ? h->sp_v.data() : i == 1 ? h->sp_p.data() : i == 2 ? h->sp_g.data() : nullptr;
}
extern "C" const long long *jac_awegpu_g_sparsity_out(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    if (!h) return nullptr;
    return i == 0 ? h->sp_jv.data() : i == 1 ? h->sp_jp.data() : nullptr;
}
extern "C" int jac_awegpu_g_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w)
{
    if (sz_arg) *sz_arg = 3;
    if (sz_res) *sz_res = 2;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}

// hess_lag for cas.nlpsol(..., {'hess_lag': external('awegpu_hess_lag', lib)}): inputs (x, p, lam_f, lam_g), output the
// upper triangle of the Hessian of lam_f f + lam_g' g (awebox/opts/default.py:323 hessian_approximation = False).  Problems
// without an objective on this path (the MPC trial, whose cost Pmpc builds itself): the constraint part, lam_f is not used.
CASADI_MEM_FUNCS(awegpu_hess_lag)
extern "C" int awegpu_hess_lag(const double **arg, double **res, long long *, double *, int mem)
{
    awegpu_handle *h = casadi_handle(mem);
    if (!h) return 1;
    if (!arg || !arg[0] || !arg[1]) {
        set_err(h, "awegpu_hess_lag: x and p must be given");
        return 1;
    }
    double *out = res ? res[0] : nullptr;
    if (!out) return 0;
    std::lock_guard<std::mutex> lock(h->casadi_mu);
    h->scratch_g.assign(h->dp.n_g, 0.0);          // NULL = all-zero multipliers
    const double zero = 0.0;
    const double *lam_f = h->hess_obj ? (arg[2] ? arg[2] : &zero) : nullptr;
    return awegpu_eval_hessian_host(h, 1, arg[0], arg[1], arg[3] ? arg[3] : h->scratch_g.data(), lam_f, 0, out);
}
extern "C" long long awegpu_hess_lag_n_in(void) { return 4; }
extern "C" long long awegpu_hess_lag_n_out(void) { return 1; }
extern "C" const char *awegpu_hess_lag_name_in(long long i) { return i == 0 ? "x" : i == 1 ? "p" : i == 2 ? "lam_f" : i == 3 ? "lam_g" : nullptr; }
extern "C" const char *awegpu_hess_lag_name_out(long long i) { return i == 0 ? "triu_hess_gamma_x_x" : nullptr; }
extern "C" const long long *awegpu_hess_lag_sparsity_in(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    if (!h || !h->hess_ready) return nullptr;
    return i == 0 ? h->sp_v.data() : i == 1 ? h->sp_p.data() : i == 2 ? h->sp_lamf.data() : i == 3 ? h->sp_g.data() : nullptr;
}
extern "C" const long long *awegpu_hess_lag_sparsity_out(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    return (h && h->hess_ready && i == 0) ? h->sp_hess.data() : nullptr;
}
extern "C" int awegpu_hess_lag_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w)
{
    if (sz_arg) *sz_arg = 4;
    if (sz_res) *sz_res = 1;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}

// objective as CasADi externals: f = external('awegpu_f', lib) (inputs V, P; output f, 1 x 1) and the Jacobian function
// external looks up by the jac_ prefix (inputs V, P, out_f; outputs jac_f_V dense 1 x n_v, jac_f_P structurally empty):
// with awegpu_g, jac_awegpu_g and awegpu_hess_lag every function nlpsol evaluates comes from this library
// (awebox/opti/preparation.py:373-379).  For an mpc trial P is Pmpc's own parameter vector (awebox/pmpc.py:152-162).
CASADI_MEM_FUNCS(awegpu_f)
CASADI_MEM_FUNCS(jac_awegpu_f)
static int casadi_eval_f(const double **arg, double **res, bool jac, int mem)
{
    awegpu_handle *h = casadi_handle(mem);
    if (!h) return 1;
    if (!arg || !arg[0] || !arg[1]) {
        set_err(h, "awegpu_f: V and P must be given");
        return 1;
    }
    double *out = res ? res[0] : nullptr;
    if (!out) return 0;
    std::lock_guard<std::mutex> lock(h->casadi_mu);
    double fval = 0.0;
    if (!jac) h->scratch_v.resize(h->dp.n_v);
    return awegpu_eval_objective_host(h, 1, arg[0], arg[1], AWEGPU_P_SHARED, jac ? &fval : out, jac ? out : h->scratch_v.data(),
                                      nullptr);
}
static const long long *sp_f_in(long long i, bool jac)
{
    awegpu_handle *h = casadi_handle(-1);
    if (!h) return nullptr;
    if (h->sp_pf.empty()) dense_sp(h->sp_pf, h->mpc_obj ? h->mpc_n_p : h->dp.n_p);
    if (h->sp_one.empty()) dense_sp(h->sp_one, 1);
    return i == 0 ? h->sp_v.data() : i == 1 ? h->sp_pf.data() : (jac && i == 2) ? h->sp_one.data() : nullptr;
}
extern "C" int awegpu_f(const double **arg, double **res, long long *, double *, int mem) { return casadi_eval_f(arg, res, false, mem); }
extern "C" long long awegpu_f_n_in(void) { return 2; }
extern "C" long long awegpu_f_n_out(void) { return 1; }
extern "C" const char *awegpu_f_name_in(long long i) { return i == 0 ? "V" : i == 1 ? "P" : nullptr; }
extern "C" const char *awegpu_f_name_out(long long i) { return i == 0 ? "f" : nullptr; }
extern "C" const long long *awegpu_f_sparsity_in(long long i) { return sp_f_in(i, false); }
extern "C" const long long *awegpu_f_sparsity_out(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    if (!h || i != 0) return nullptr;
    if (h->sp_one.empty()) dense_sp(h->sp_one, 1);
    return h->sp_one.data();
}
extern "C" int awegpu_f_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w)
{
    if (sz_arg) *sz_arg = 2;
    if (sz_res) *sz_res = 1;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}
extern "C" int jac_awegpu_f(const double **arg, double **res, long long *, double *, int mem) { return casadi_eval_f(arg, res, true, mem); }
extern "C" long long jac_awegpu_f_n_in(void) { return 3; }
extern "C" long long jac_awegpu_f_n_out(void) { return 2; }
extern "C" const char *jac_awegpu_f_name_in(long long i) { return i == 0 ? "V" : i == 1 ? "P" : i == 2 ? "out_f" : nullptr; }
extern "C" const char *jac_awegpu_f_name_out(long long i) { return i == 0 ? "jac_f_V" : i == 1 ? "jac_f_P" : nullptr; }
extern "C" const long long *jac_awegpu_f_sparsity_in(long long i) { return sp_f_in(i, true); }
extern "C" const long long *jac_awegpu_f_sparsity_out(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    if (!h) return nullptr;
    if (h->sp_jfv.empty()) {
        // 1 x n_v dense in compact CCS: [nrow, ncol, colind[ncol + 1], row[nnz]]
        const long long n = h->dp.n_v;
        h->sp_jfv.assign(2 + (size_t)n + 1 + (size_t)n, 0);
        h->sp_jfv[0] = 1, h->sp_jfv[1] = n;
        for (long long c = 0; c <= n; ++c) h->sp_jfv[2 + c] = c;
        const long long np_ = h->mpc_obj ? h->mpc_n_p : h->dp.n_p;
        h->sp_jfp.assign(2 + (size_t)np_ + 1, 0);      // 1 x n_p, no non-zeros (not a source of parametric sensitivities)
        h->sp_jfp[0] = 1, h->sp_jfp[1] = np_;
    }
    return i == 0 ? h->sp_jfv.data() : i == 1 ? h->sp_jfp.data() : nullptr;
}
extern "C" int jac_awegpu_f_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w)
{
    if (sz_arg) *sz_arg = 3;
    if (sz_res) *sz_res = 2;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}

// outputs map as a CasADi external: Outputs_fun = external('awegpu_outputs', lib) (awebox/ocp/nlp.py:335-336): inputs V, P;
// output the dense n_out x n_nodes matrix in CasADi's column-major non-zero order (node after node), rows in the order
// of awegpu_outputs_label
CASADI_MEM_FUNCS(awegpu_outputs)
extern "C" int awegpu_outputs(const double **arg, double **res, long long *, double *, int mem)
{
    awegpu_handle *h = casadi_handle(mem);
    if (!h) return 1;
    if (!arg || !arg[0] || !arg[1]) {
        set_err(h, "awegpu_outputs: V and P must be given");
        return 1;
    }
    double *out = res ? res[0] : nullptr;
    if (!out) return 0;
    std::lock_guard<std::mutex> lock(h->casadi_mu);
    const size_t n_out = (size_t)awe_out_n(), n_nodes = (size_t)h->dp.n_k * (N_D + 1);
    h->scratch_o.resize(n_out * n_nodes);
    const int rc = awegpu_eval_outputs_host(h, 1, arg[0], arg[1], AWEGPU_P_SHARED, h->scratch_o.data());
    if (rc) return rc;
    for (size_t n = 0; n < n_nodes; ++n)
        for (size_t i = 0; i < n_out; ++i) out[n * n_out + i] = h->scratch_o[i * n_nodes + n];
    return 0;
}
extern "C" long long awegpu_outputs_n_in(void) { return 2; }
extern "C" long long awegpu_outputs_n_out(void) { return 1; }
extern "C" const char *awegpu_outputs_name_in(long long i) { return i == 0 ? "V" : i == 1 ? "P" : nullptr; }
extern "C" const char *awegpu_outputs_name_out(long long i) { return i == 0 ? "outputs" : nullptr; }
extern "C" const long long *awegpu_outputs_sparsity_in(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    if (!h) return nullptr;
    return i == 0 ? h->sp_v.data() : i == 1 ? h->sp_p.data() : nullptr;
}
extern "C" const long long *awegpu_outputs_sparsity_out(long long i)
{
    awegpu_handle *h = casadi_handle(-1);
    if (!h || i != 0) return nullptr;
    h->sp_out = {(long long)awe_out_n(), (long long)h->dp.n_k * (N_D + 1), 1};      // compact dense flag
    return h->sp_out.data();
}
extern "C" int awegpu_outputs_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w)
{
    if (sz_arg) *sz_arg = 2;
    if (sz_res) *sz_res = 1;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}
#undef CASADI_MEM_FUNCS

#endif  /* TU_OWNS(0) */
