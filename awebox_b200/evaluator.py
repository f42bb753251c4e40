"""Host-side mirror of the reference interface for the hot path.

``Evaluator`` plays the role of the reference's ``NLP.g_fun`` (a ``casadi.Function`` named
'cstr_fun' with inputs ``(V, P)`` and one output ``g``; awebox/ocp/nlp.py:119-121, 270-272) and of the
``jac_g`` function CasADi derives from it for IPOPT (awebox/opti/preparation.py:379): same argument
meaning and order, same row / CCS ordering, evaluated for a whole batch of instances per call on the
GPU through the C ABI (include/awegpu.h).  There is no CPU fallback: construction raises when the
compiled CUDA library is missing or no CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

from . import cabi
from .build import lib_path
from .ocp.nlp import OcpProblem
from .spec import ModelSpec


class AwegpuError(RuntimeError):
    pass


def load_library(spec_name):
    path = os.environ.get('AWEGPU_LIB') or lib_path(spec_name)      # AWEGPU_LIB: kernel-tuning experiments
    if not os.path.exists(path):
        raise AwegpuError('CUDA library %s is missing: run `python -m awebox_b200.build %s` '
                          '(there is no CPU fallback)' % (path, spec_name))
    lib = C.CDLL(path)
    lib.awegpu_model_name.restype = C.c_char_p
    lib.awegpu_last_error.restype = C.c_char_p
    lib.awegpu_last_error.argtypes = [C.c_void_p]
    lib.awegpu_create.restype = C.c_int
    lib.awegpu_create.argtypes = [C.POINTER(cabi.ProblemStruct), C.c_int, C.POINTER(C.c_void_p)]
    lib.awegpu_destroy.argtypes = [C.c_void_p]
    lib.awegpu_destroy.restype = None
    lib.awegpu_sparsity.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.awegpu_eval_batch.restype = C.c_int
    lib.awegpu_eval_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_void_p]
    lib.awegpu_eval_batch_host.restype = C.c_int
    lib.awegpu_eval_batch_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int32]
    lib.awegpu_reduce_batch.restype = C.c_int
    lib.awegpu_reduce_batch.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 6
    lib.awegpu_eval_batch_reduce_host.restype = C.c_int
    lib.awegpu_eval_batch_reduce_host.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32, C.c_int32] + \
        [C.c_void_p] * 6
    lib.awegpu_eval_objective.restype = C.c_int
    lib.awegpu_eval_objective.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 4
    lib.awegpu_eval_objective_host.restype = C.c_int
    lib.awegpu_eval_objective_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 3
    lib.awegpu_last_eval_ms.restype = C.c_float
    lib.awegpu_last_eval_ms.argtypes = [C.c_void_p]
    lib.awegpu_last_kernel_ms.restype = C.c_float
    lib.awegpu_last_kernel_ms.argtypes = [C.c_void_p]
    lib.awegpu_kernel_info.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.awegpu_launch_count.restype = C.c_int64
    lib.awegpu_launch_count.argtypes = [C.c_void_p]
    lib.awegpu_synchronize.argtypes = [C.c_void_p]
    lib.awegpu_casadi_bind.argtypes = [C.c_void_p]
    lib.awegpu_hess_setup.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64] + [C.c_void_p] * 5 + [C.c_int32] * 3 + \
        [C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p]
    lib.awegpu_hess_nnz.argtypes = [C.c_void_p]
    lib.awegpu_hess_nnz.restype = C.c_int64
    lib.awegpu_hess_sparsity.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.awegpu_eval_hessian.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p, C.c_void_p]
    lib.awegpu_eval_hessian_host.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p]
    lib.awegpu_last_hessian_ms.argtypes = [C.c_void_p]
    lib.awegpu_last_hessian_ms.restype = C.c_float
    lib.awegpu_model_dims.argtypes = [C.POINTER(C.c_int32)] * 4 + [C.POINTER(C.c_int64)] * 2
    lib.awegpu_outputs_n.restype = C.c_int32
    lib.awegpu_outputs_label.restype = C.c_char_p
    lib.awegpu_outputs_label.argtypes = [C.c_int32]
    lib.awegpu_outputs_nodes.restype = C.c_int32
    lib.awegpu_outputs_nodes.argtypes = [C.c_void_p]
    lib.awegpu_eval_outputs.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.awegpu_eval_outputs_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.awegpu_last_outputs_ms.argtypes = [C.c_void_p]
    lib.awegpu_last_outputs_ms.restype = C.c_float
    if lib.awegpu_model_name().decode() != spec_name:
        raise AwegpuError('library %s was generated for model %r' % (path, lib.awegpu_model_name().decode()))
    return lib


class Evaluator(object):
    def __init__(self, spec, n_k=None, device=0, ocp=None):
        if isinstance(spec, str):
            spec = ModelSpec.load(spec)
        self.spec = spec
        self.ocp = ocp if ocp is not None else OcpProblem(spec, n_k=n_k)
        self.lib = load_library(spec.name)
        self.problem, self._keep = cabi.make_problem(self.ocp)
        self.handle = C.c_void_p()
        rc = self.lib.awegpu_create(C.byref(self.problem), int(device), C.byref(self.handle))
        if rc != 0:
            raise AwegpuError('awegpu_create failed (%d): %s' % (rc, self.lib.awegpu_last_error(None).decode()))
        self.device = int(device)
        self.n_v, self.n_p, self.n_g, self.nnz = self.ocp.n_v, self.ocp.n_p, self.ocp.n_g, self.ocp.nnz
        self.n_theta0 = self.ocp.n_theta0
        # width of the parameter vectors of eval_objective: the NLP's P, or Pmpc's own parameters for an MPC trial
        self.n_p_objective = int(self.problem.mpc_n_p) if int(self.problem.obj_enabled) == 2 else self.n_p
        fc, fs = C.c_int64(0), C.c_int64(0)
        self.lib.awegpu_model_dims(None, None, None, None, C.byref(fc), C.byref(fs))
        self.flops_coll, self.flops_shoot = int(fc.value), int(fs.value)

    def close(self):
        if getattr(self, 'handle', None) and self.handle.value:
            self.lib.awegpu_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise AwegpuError('%s failed (%d): %s' % (what, rc, self.lib.awegpu_last_error(self.handle).decode()))

    # ---- structure (bit-exact integer data)
    def sparsity(self):
        """CCS pattern of jac_g w.r.t. V as (colind [n_v+1], row [nnz]) int64 — casadi_int."""
        colind = np.empty(self.n_v + 1, dtype=np.int64)
        row = np.empty(self.nnz, dtype=np.int64)
        self.lib.awegpu_sparsity(self.handle, colind.ctypes.data, row.ctypes.data)
        return colind, row

    @staticmethod
    def p_flags(P, n_p, n_theta0, B):
        shape = tuple(P.shape)
        flags = 0
        if shape[-1] == n_theta0:
            flags |= cabi.AWEGPU_P_THETA0
        elif shape[-1] != n_p:
            raise ValueError('P has %d columns; expected n_p=%d or n_theta0=%d' % (shape[-1], n_p, n_theta0))
        if len(shape) == 1 or shape[0] == 1:
            flags |= cabi.AWEGPU_P_SHARED
        elif shape[0] != B:
            raise ValueError('P has %d rows for a batch of %d' % (shape[0], B))
        return flags

    # ---- host entry: numpy in, numpy out (H2D / D2H inside the call)
    def eval(self, W, P, jac=True, batch_minor=False):
        """g and jac_g values for a batch.  W [B, n_v] (or [n_v]); P [B, n_p] | [n_p] | [B, n_theta0] |
        [n_theta0].  Returns (G [B, n_g], Jv [B, nnz]) — Jv in CCS order of ``sparsity()``
        (``[nnz, B]`` / ``[n_g, B]`` when batch_minor)."""
        W = np.ascontiguousarray(W, dtype=np.float64)
        if W.ndim == 1:
            W = W[None, :]
        if W.shape[1] != self.n_v:
            raise ValueError('W has %d columns, expected n_v=%d' % (W.shape[1], self.n_v))
        B = W.shape[0]
        P = np.ascontiguousarray(P, dtype=np.float64)
        flags = self.p_flags(P, self.n_p, self.n_theta0, B)
        if batch_minor:
            flags |= cabi.AWEGPU_JAC_BATCH_MINOR
        if not jac:
            flags |= cabi.AWEGPU_SKIP_JAC
        G = np.empty((self.n_g, B) if batch_minor else (B, self.n_g))
        Jv = (np.empty((self.nnz, B) if batch_minor else (B, self.nnz))) if jac else None
        rc = self.lib.awegpu_eval_batch_host(self.handle, B, W.ctypes.data, P.ctypes.data, G.ctypes.data,
                                             Jv.ctypes.data if jac else None, flags)
        self._check(rc, 'awegpu_eval_batch_host')
        return G, Jv

    # ---- device entry: torch CUDA tensors (device memory stays with the caller)
    def eval_device(self, W, P, G, Jv=None, batch_minor=False, stream=None):
        import torch
        for t in (W, P, G) + ((Jv,) if Jv is not None else ()):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise ValueError('eval_device needs contiguous float64 CUDA tensors')
        B = W.shape[0]
        flags = self.p_flags(P, self.n_p, self.n_theta0, B)
        if batch_minor:
            flags |= cabi.AWEGPU_JAC_BATCH_MINOR
        if Jv is None:
            flags |= cabi.AWEGPU_SKIP_JAC
        if stream is None:
            stream = torch.cuda.current_stream(W.device).cuda_stream
        if not stream:
            # torch's default stream is the legacy NULL stream; the C ABI reads NULL as "the handle's own
            # stream", so name the legacy stream explicitly (cudaStreamLegacy) to stay ordered with torch
            stream = 1
        rc = self.lib.awegpu_eval_batch(self.handle, B, W.data_ptr(), P.data_ptr(), G.data_ptr(),
                                        Jv.data_ptr() if Jv is not None else None, flags, C.c_void_p(stream))
        self._check(rc, 'awegpu_eval_batch')

    # ---- device-resident products for a batched consumer
    def reduce_device(self, G, Jv, lam=None, v=None, want=('JTl', 'Jv', 'cnorm', 'rnorm', 'ginf'), stream=None):
        """products of the Jacobians a preceding ``eval_device`` left on the device (torch CUDA tensors):
        ``JTl = J^T lam`` [B, n_v], ``Jv = J v`` [B, n_g], column / row max-norms, ``ginf = max |g|`` [B].
        ``lam`` [B, n_g] or [n_g], ``v`` [B, n_v] or [n_v].  -> dict of torch tensors."""
        import torch
        B = int(G.shape[0])
        red = 0
        if lam is not None and lam.dim() == 1:
            red |= cabi.AWEGPU_RED_LAMBDA_SHARED
        if v is not None and v.dim() == 1:
            red |= cabi.AWEGPU_RED_V_SHARED
        out = {}
        shapes = {'JTl': (B, self.n_v), 'Jv': (B, self.n_g), 'cnorm': (B, self.n_v), 'rnorm': (B, self.n_g), 'ginf': (B,)}
        for key in want:
            if (key == 'JTl' and lam is None) or (key == 'Jv' and v is None):
                continue
            out[key] = torch.empty(shapes[key], dtype=torch.float64, device=G.device)
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        st = stream.cuda_stream if stream is not None else torch.cuda.current_stream(G.device).cuda_stream
        st = C.c_void_p(st if st else 1)           # legacy NULL stream by name, as in eval_device
        rc = self.lib.awegpu_reduce_batch(self.handle, B, ptr(G), ptr(Jv), ptr(lam), ptr(v), red, ptr(out.get('JTl')),
                                          ptr(out.get('Jv')), ptr(out.get('cnorm')), ptr(out.get('rnorm')),
                                          ptr(out.get('ginf')), st)
        self._check(rc, 'awegpu_reduce_batch')
        return out

    def eval_reduce(self, W, P, lam=None, v=None, want=('JTl', 'Jv', 'cnorm', 'rnorm', 'ginf')):
        """host entry: g and the products (numpy); the Jacobian values never leave the device."""
        W = np.ascontiguousarray(np.atleast_2d(np.asarray(W, dtype=np.float64)))
        B = W.shape[0]
        P = np.ascontiguousarray(np.asarray(P, dtype=np.float64))
        flags = self.p_flags(P, self.n_p, self.n_theta0, B)
        red = 0
        if lam is not None:
            lam = np.ascontiguousarray(np.asarray(lam, dtype=np.float64))
            red |= cabi.AWEGPU_RED_LAMBDA_SHARED if lam.ndim == 1 else 0
        if v is not None:
            v = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
            red |= cabi.AWEGPU_RED_V_SHARED if v.ndim == 1 else 0
        G = np.empty((B, self.n_g))
        shapes = {'JTl': (B, self.n_v), 'Jv': (B, self.n_g), 'cnorm': (B, self.n_v), 'rnorm': (B, self.n_g), 'ginf': (B,)}
        out = {}
        for key in want:
            if (key == 'JTl' and lam is None) or (key == 'Jv' and v is None):
                continue
            out[key] = np.empty(shapes[key])
        ptr = lambda a: a.ctypes.data if a is not None else None
        rc = self.lib.awegpu_eval_batch_reduce_host(self.handle, B, ptr(W), ptr(P), ptr(lam), ptr(v), flags, red, ptr(G),
                                                    ptr(out.get('JTl')), ptr(out.get('Jv')), ptr(out.get('cnorm')),
                                                    ptr(out.get('rnorm')), ptr(out.get('ginf')))
        self._check(rc, 'awegpu_eval_batch_reduce_host')
        out['G'] = G
        return out

    # ---- objective (the role of the reference's NLP.f_fun and of the gradient nlpsol derives from it)
    COMPONENTS = ('tracking', 'xdot_regularisation', 'u_regularisation', 'fictitious', 'theta_regularisation', 'beta', 'time',
                  'power', 'homotopy', 'general_problem')

    def eval_objective(self, W, P, components=False):
        """f [B] and grad f [B, n_v] (dense) for a batch; P: full parameter vectors [B, n_p] or one [n_p]
        (the objective reads p.ref, p.weights, the cost factors and theta0); for an MPC trial: Pmpc's own parameter
        vectors [B, n_p_objective] (x0, ref, u_ref, Q, R, P).  numpy in, numpy out."""
        W = np.ascontiguousarray(np.atleast_2d(np.asarray(W, dtype=np.float64)))
        B = W.shape[0]
        P = np.ascontiguousarray(np.asarray(P, dtype=np.float64))
        flags = self.p_flags(P, self.n_p_objective, -1, B)
        F, Gr = np.empty(B), np.empty((B, self.n_v))
        comp = np.empty((B, len(self.COMPONENTS))) if components else None
        rc = self.lib.awegpu_eval_objective_host(self.handle, B, W.ctypes.data, P.ctypes.data, flags, F.ctypes.data,
                                                 Gr.ctypes.data, comp.ctypes.data if components else None)
        self._check(rc, 'awegpu_eval_objective_host')
        return (F, Gr, comp) if components else (F, Gr)

    def eval_objective_device(self, W, P, F, GradF, stream=None):
        """device entry (torch CUDA tensors, float64, contiguous)"""
        import torch
        B = W.shape[0]
        flags = self.p_flags(P, self.n_p_objective, -1, B)
        st = stream if stream is not None else torch.cuda.current_stream(W.device).cuda_stream
        rc = self.lib.awegpu_eval_objective(self.handle, B, W.data_ptr(), P.data_ptr(), flags, F.data_ptr(), GradF.data_ptr(),
                                            None, C.c_void_p(st if st else 1))
        self._check(rc, 'awegpu_eval_objective')

    def f_fun(self, V, P):
        """f(V, P) for one instance (the reference's ``nlp.f_fun``)."""
        return float(self.eval_objective(np.asarray(V, dtype=float).reshape(1, -1), np.asarray(P, dtype=float).reshape(-1))[0][0])

    def grad_f_fun(self, V, P):
        return self.eval_objective(np.asarray(V, dtype=float).reshape(1, -1), np.asarray(P, dtype=float).reshape(-1))[1][0]

    # ---- second-order hook: triu of d^2/dw^2 (lam_f f + lam_g' g)
    def hessian_setup(self):
        """derive the node Hessians' tables (awebox_b200/ocp/hessian.py) and upload them; idempotent"""
        if getattr(self, 'hess', None) is not None:
            return self.hess
        from .ocp.hessian import LagrangianHessian, has_objective
        ocp = self.ocp
        with_obj = has_objective(ocp)
        if bool(self.lib.awegpu_hess_has_objective()) != with_obj:
            raise AwegpuError('the library was generated %s the objective part of the Hessian' % ('without' if with_obj else 'with'))
        H = LagrangianHessian(ocp, objective=with_obj)
        t_terms = H.T_terms if with_obj else []
        keep = [np.ascontiguousarray(a) for a in (H.colind.astype(np.int64), H.row.astype(np.int64), H.asm_ptr.astype(np.int64),
                                                  H.asm_idx.astype(np.int64), H.asm_coef.astype(np.float64),
                                                  np.asarray(ocp.coll.coeff_collocation, dtype=np.float64),
                                                  np.asarray(ocp.coll.quad_weights, dtype=np.float64).ravel(),
                                                  np.array([i for i, _ in t_terms] + [0, 0], dtype=np.int32),
                                                  np.array([c for _, c in t_terms] + [0.0, 0.0], dtype=np.float64))]
        row_stride = ocp.row_shoot[1] - ocp.row_shoot[0] if ocp.n_k > 1 else ocp.n_shoot_eq + ocp.n_ineq + ocp.d * ocp.n_eq + ocp.n_x
        rc = self.lib.awegpu_hess_setup(self.handle, H.nnz, H.n_hv, len(H.asm_idx), keep[0].ctypes.data, keep[1].ctypes.data,
                                        keep[2].ctypes.data, keep[3].ctypes.data, keep[4].ctypes.data, int(ocp.v_x[0]),
                                        int(ocp.row_shoot[0]), int(row_stride), keep[5].ctypes.data, keep[6].ctypes.data,
                                        H.n_ec, H.n_es, int(with_obj), len(t_terms), keep[7].ctypes.data, keep[8].ctypes.data)
        self._check(rc, 'awegpu_hess_setup')
        self.hess, self.nnz_hess, self.hess_objective = H, H.nnz, with_obj
        return H

    def hess_sparsity(self):
        """CCS pattern of triu(hess_lag) w.r.t. V as (colind [n_v+1], row [nnz_hess]) int64 -- casadi_int"""
        self.hessian_setup()
        colind = np.empty(self.n_v + 1, dtype=np.int64)
        row = np.empty(self.nnz_hess, dtype=np.int64)
        self._check(self.lib.awegpu_hess_sparsity(self.handle, colind.ctypes.data, row.ctypes.data), 'awegpu_hess_sparsity')
        return colind, row

    def eval_hessian(self, W, P, Lam, sigma=None):
        """non-zeros of triu(d^2/dw^2 (sigma f + lam' g)) for a batch: W [B, n_v], P as ``eval``, Lam [B, n_g], sigma [B] (or a
        scalar; None: the constraint part alone) -> [B, nnz_hess]"""
        self.hessian_setup()
        W = np.ascontiguousarray(W, dtype=np.float64)
        if W.ndim == 1:
            W = W[None, :]
        B = W.shape[0]
        Lam = np.ascontiguousarray(Lam, dtype=np.float64).reshape(B, self.n_g)
        P = np.ascontiguousarray(P, dtype=np.float64)
        flags = self.p_flags(P, self.n_p, self.n_theta0, B)
        if sigma is not None:
            sigma = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, dtype=np.float64), (B,)))
        H = np.empty((B, self.nnz_hess))
        rc = self.lib.awegpu_eval_hessian_host(self.handle, B, W.ctypes.data, P.ctypes.data, Lam.ctypes.data,
                                               sigma.ctypes.data if sigma is not None else None, flags, H.ctypes.data)
        self._check(rc, 'awegpu_eval_hessian_host')
        return H

    def eval_hessian_device(self, W, P, Lam, H, sigma=None, stream=None):
        import torch
        self.hessian_setup()
        for t in (W, P, Lam, H) + ((sigma,) if sigma is not None else ()):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise ValueError('eval_hessian_device needs contiguous float64 CUDA tensors')
        B = W.shape[0]
        flags = self.p_flags(P, self.n_p, self.n_theta0, B)
        if stream is None:
            stream = torch.cuda.current_stream(W.device).cuda_stream
        if not stream:
            stream = 1
        rc = self.lib.awegpu_eval_hessian(self.handle, B, W.data_ptr(), P.data_ptr(), Lam.data_ptr(),
                                          sigma.data_ptr() if sigma is not None else None, flags, H.data_ptr(), C.c_void_p(stream))
        self._check(rc, 'awegpu_eval_hessian')

    def last_hessian_ms(self):
        return float(self.lib.awegpu_last_hessian_ms(self.handle))

    def hess_lag_fun(self, V, P, lam_f, lam_g):
        """one instance, the way nlpsol calls hess_lag(x, p, lam_f, lam_g): values of triu(H) in CCS order"""
        self.hessian_setup()
        return self.eval_hessian(np.asarray(V, dtype=np.float64)[None, :], np.asarray(P, dtype=np.float64),
                                 np.asarray(lam_g, dtype=np.float64)[None, :],
                                 sigma=float(lam_f) if self.hess_objective else None)[0]

    # ------------------------------------------------------------------ outputs map
    def outputs_labels(self):
        """labels '[group,name,i]' of the rows of the outputs map, as in ``model.outputs.labels()``"""
        return [self.lib.awegpu_outputs_label(i).decode() for i in range(int(self.lib.awegpu_outputs_n()))]

    def outputs_index(self, group, name=None):
        pre = '[%s,' % group if name is None else '[%s,%s,' % (group, name)
        return [i for i, l in enumerate(self.outputs_labels()) if l.startswith(pre)]

    def eval_outputs(self, W, P):
        """``nlp.Outputs_fun(V, P)`` for a batch: W [B, n_v], P as ``eval`` -> [B, n_out, n_k (d+1)] (nodes of interval
        k: shooting node, then the d collocation nodes)"""
        W = np.ascontiguousarray(W, dtype=np.float64)
        if W.ndim == 1:
            W = W[None, :]
        B = W.shape[0]
        P = np.ascontiguousarray(P, dtype=np.float64)
        flags = self.p_flags(P, self.n_p, self.n_theta0, B)
        out = np.empty((B, int(self.lib.awegpu_outputs_n()), int(self.lib.awegpu_outputs_nodes(self.handle))))
        rc = self.lib.awegpu_eval_outputs_host(self.handle, B, W.ctypes.data, P.ctypes.data, flags, out.ctypes.data)
        self._check(rc, 'awegpu_eval_outputs_host')
        return out

    def eval_outputs_device(self, W, P, Out, stream=None):
        import torch
        for t in (W, P, Out):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise ValueError('eval_outputs_device needs contiguous float64 CUDA tensors')
        B = W.shape[0]
        flags = self.p_flags(P, self.n_p, self.n_theta0, B)
        if stream is None:
            stream = torch.cuda.current_stream(W.device).cuda_stream
        if not stream:
            stream = 1
        rc = self.lib.awegpu_eval_outputs(self.handle, B, W.data_ptr(), P.data_ptr(), flags, Out.data_ptr(), C.c_void_p(stream))
        self._check(rc, 'awegpu_eval_outputs')

    def last_outputs_ms(self):
        return float(self.lib.awegpu_last_outputs_ms(self.handle))

    def outputs_fun(self, V, P):
        """one instance, the way the reference calls ``nlp.Outputs_fun(V, P)``: -> [n_out, n_nodes]"""
        return self.eval_outputs(np.asarray(V, dtype=np.float64)[None, :], np.asarray(P, dtype=np.float64))[0]

    def last_eval_ms(self):
        return float(self.lib.awegpu_last_eval_ms(self.handle))

    def last_kernel_ms(self):
        """device time of the dominant kernel (collocation node programs, output stage) of the last eval."""
        return float(self.lib.awegpu_last_kernel_ms(self.handle))

    def kernel_info(self):
        g, e = (C.c_int32 * 6)(), (C.c_int32 * 2)()
        oc, os_ = C.c_int64(0), C.c_int64(0)
        self.lib.awegpu_kernel_info(g, e, C.byref(oc), C.byref(os_))
        return {'groups_coll': list(g[:3]), 'groups_shoot': list(g[3:]), 'exports_coll': int(e[0]),
                'exports_shoot': int(e[1]), 'ops_coll_node': int(oc.value), 'ops_shoot_node': int(os_.value)}

    def launch_count(self):
        return int(self.lib.awegpu_launch_count(self.handle))

    # ---- single-instance, reference-style callables
    def g_fun(self, V, P):
        """``nlp.g_fun(V, P)``: one instance, returns g [n_g]."""
        G, _ = self.eval(np.asarray(V, dtype=float).reshape(1, -1), np.asarray(P, dtype=float).reshape(-1), jac=False)
        return G[0]

    def jac_g_fun(self, V, P):
        """values of jac_g(V, P) in CCS order, [nnz]."""
        _, Jv = self.eval(np.asarray(V, dtype=float).reshape(1, -1), np.asarray(P, dtype=float).reshape(-1))
        return Jv[0]

    def jac_g_scipy(self, V, P):
        import scipy.sparse as sps
        colind, row = self.sparsity()
        return sps.csc_matrix((self.jac_g_fun(V, P), row, colind), shape=(self.n_g, self.n_v))
