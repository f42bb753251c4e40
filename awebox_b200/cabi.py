"""ctypes mirror of include/awegpu.h (``awegpu_problem_t``) and its construction from an OcpProblem."""
import ctypes as C

import numpy as np

P_I32 = C.POINTER(C.c_int32)
P_I64 = C.POINTER(C.c_int64)
P_F64 = C.POINTER(C.c_double)

AWEGPU_JAC_INSTANCE_MAJOR = 0
AWEGPU_JAC_BATCH_MINOR = 1
AWEGPU_P_FULL = 0
AWEGPU_P_THETA0 = 4
AWEGPU_P_SHARED = 8
AWEGPU_SKIP_JAC = 16
AWEGPU_HOST_PINNED = 32
AWEGPU_RED_LAMBDA_SHARED = 1
AWEGPU_RED_V_SHARED = 2


class ProblemStruct(C.Structure):
    _fields_ = [
        ('n_v', C.c_int32), ('n_p', C.c_int32), ('n_g', C.c_int32),
        ('nnz', C.c_int64),
        ('n_k', C.c_int32), ('d', C.c_int32),
        ('n_x', C.c_int32), ('n_u', C.c_int32), ('n_z', C.c_int32), ('n_theta', C.c_int32),
        ('n_phi', C.c_int32), ('n_theta0', C.c_int32), ('n_var', C.c_int32),
        ('v_phi', C.c_int32), ('p_theta0', C.c_int32),
        ('n_coll_out', C.c_int32), ('n_shoot_out', C.c_int32),
        ('n_shoot_eq', C.c_int32), ('n_ineq', C.c_int32),
        ('n_const', C.c_int32), ('n_periodic', C.c_int32), ('row_periodic', C.c_int32), ('x_last', C.c_int32),
        ('x_base', P_I32), ('x_next', P_I32),
        ('u_base', P_I32), ('xdot_base', P_I32), ('z_base', P_I32),
        ('cz_base', P_I32), ('theta_idx', P_I32), ('tf_idx', P_I32),
        ('row_shoot', P_I32), ('row_path', P_I32), ('row_cont', P_I32), ('row_coll', P_I32),
        ('periodic_idx', P_I32),
        ('coll_slot', P_I32), ('shoot_slot', P_I32), ('const_slot', P_I32),
        ('const_val', P_F64), ('const_begin', P_I32),
        ('coeff_coll', P_F64), ('coeff_cont', P_F64),
        ('colind', P_I64), ('row', P_I64),
        ('n_aff', C.c_int32),
        ('aff_row', P_I32), ('aff_begin', P_I32), ('aff_col', P_I32), ('aff_coef', P_F64),
        ('affp_begin', P_I32), ('affp_col', P_I32), ('affp_coef', P_F64), ('aff_const', P_F64),
        ('constraint_scale', C.c_double),
        ('n_kcoll_out', C.c_int32), ('n_kcoll_head', C.c_int32), ('kcoll_slot', P_I32),
        ('obj_enabled', C.c_int32), ('obj_cat', P_I32), ('obj_norm', P_F64), ('obj_quad', P_F64),
        ('obj_n_tf', C.c_int32), ('obj_tf_idx', P_I32), ('obj_tf_coef', P_F64),
        ('obj_p_weights', C.c_int32), ('obj_p_cost', C.c_int32), ('obj_psi_idx', C.c_int32), ('obj_phi_cost', P_I32),
        ('obj_c_power', C.c_int32), ('obj_c_tf', C.c_int32), ('obj_c_beta', C.c_int32), ('obj_c_cat', P_I32),
        ('mpc_n_p', C.c_int32), ('mpc_p_ref', C.c_int32), ('mpc_p_Q', C.c_int32), ('mpc_p_R', C.c_int32),
        ('mpc_p_P', C.c_int32), ('mpc_kind', P_I32), ('mpc_idx', P_I32), ('mpc_fac', P_F64),
    ]


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64).reshape(-1))


def make_problem(ocp):
    """-> (ProblemStruct, keepalive list of numpy arrays)."""
    n_k, d = ocp.n_k, ocp.d
    arrays = {
        'x_base': _i32([[ocp.x_base(k, r) for r in range(d + 1)] for k in range(n_k)]),
        'x_next': _i32([ocp.v_x[k + 1] for k in range(n_k)]),
        'u_base': _i32(ocp.v_u), 'xdot_base': _i32(ocp.v_xdot), 'z_base': _i32(ocp.v_z),
        'cz_base': _i32(ocp.v_cz), 'theta_idx': _i32(ocp.theta_idx), 'tf_idx': _i32(ocp.tf_idx),
        'row_shoot': _i32(ocp.row_shoot), 'row_path': _i32(ocp.row_path), 'row_cont': _i32(ocp.row_cont),
        'row_coll': _i32(ocp.row_coll), 'periodic_idx': _i32(ocp.periodic_idx),
        'coll_slot': _i32(ocp.coll_slot), 'shoot_slot': _i32(ocp.shoot_slot),
        'const_slot': _i32(ocp.const_slot), 'const_val': _f64(ocp.const_val),
        'const_begin': _i32(ocp.const_begin),
        'coeff_coll': _f64(ocp.coll.coeff_collocation), 'coeff_cont': _f64(ocp.coll.coeff_continuity),
        'colind': _i64(ocp.colind), 'row': _i64(ocp.row),
        'aff_row': _i32(ocp.aff_row), 'aff_begin': _i32(ocp.aff_begin), 'aff_col': _i32(ocp.aff_col),
        'aff_coef': _f64(ocp.aff_coef), 'affp_begin': _i32(ocp.affp_begin), 'affp_col': _i32(ocp.affp_col),
        'affp_coef': _f64(ocp.affp_coef), 'aff_const': _f64(ocp.aff_const),
        'kcoll_slot': _i32(ocp.kcoll_slot),
    }
    ps = ProblemStruct()
    ps.n_v, ps.n_p, ps.n_g, ps.nnz = ocp.n_v, ocp.n_p, ocp.n_g, ocp.nnz
    ps.n_k, ps.d = n_k, d
    ps.n_x, ps.n_u, ps.n_z, ps.n_theta = ocp.n_x, ocp.n_u, ocp.n_z, ocp.n_theta
    ps.n_phi, ps.n_theta0, ps.n_var = ocp.n_phi, ocp.n_theta0, ocp.n_var
    ps.v_phi, ps.p_theta0 = ocp.v_phi, ocp.p_theta0
    ps.n_coll_out, ps.n_shoot_out = len(ocp.coll_out), len(ocp.shoot_out)
    ps.n_shoot_eq, ps.n_ineq = ocp.n_shoot_eq, ocp.n_ineq
    ps.n_const, ps.n_periodic = len(ocp.const_slot), len(ocp.periodic_idx)
    ps.row_periodic, ps.x_last = ocp.row_periodic, ocp.v_cx[n_k - 1][d - 1]
    ps.constraint_scale = ocp.constraint_scale
    ps.n_aff = ocp.n_aff
    ps.n_kcoll_out, ps.n_kcoll_head = len(ocp.kcoll_out), ocp.n_kcoll_head
    # objective (power-cycle trajectories only)
    ps.obj_enabled = 0
    try:
        from .ocp.objective import Objective, REG_CATEGORIES, COST_NAMES
        obj = Objective(ocp)
    except NotImplementedError:
        obj = None
    if obj is not None:
        terms = obj.time_period_terms()
        arrays.update({
            'obj_cat': _i32(obj.cat_index), 'obj_norm': _f64([obj.norm[c] for c in REG_CATEGORIES + ['beta']]),
            'obj_quad': _f64(obj.quad), 'obj_tf_idx': _i32([i for i, _ in terms]), 'obj_tf_coef': _f64([c for _, c in terms]),
            'obj_phi_cost': _i32([COST_NAMES.index(n) for n in obj.phi_names]),
            'obj_c_cat': _i32([COST_NAMES.index(c) for c in REG_CATEGORIES]),
        })
        ps.obj_enabled, ps.obj_n_tf = 1, len(terms)
        ps.obj_p_weights, ps.obj_p_cost = obj.p_weights, obj.p_cost
        ps.obj_psi_idx = obj.phi_idx[obj.phi_names.index('psi')]
        ps.obj_c_power, ps.obj_c_tf, ps.obj_c_beta = COST_NAMES.index('power'), COST_NAMES.index('t_f'), COST_NAMES.index('beta')
    if ocp.traj_type == 'mpc':
        from .ocp.objective import MpcObjective
        mo = MpcObjective(ocp)
        arrays.update({'mpc_kind': _i32(mo.kind), 'mpc_idx': _i32(mo.idx), 'mpc_fac': _f64(mo.fac)})
        ps.obj_enabled = 2
        ps.mpc_n_p, ps.mpc_p_ref, ps.mpc_p_Q, ps.mpc_p_R, ps.mpc_p_P = mo.n_p, mo.p_ref, mo.p_Q, mo.p_R, mo.p_P
    for name, arr in arrays.items():
        ctype = dict(ProblemStruct._fields_)[name]
        setattr(ps, name, arr.ctypes.data_as(ctype))
    return ps, arrays
