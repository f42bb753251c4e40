"""Outputs map of the OCP: model.outputs_fun over the shooting and collocation nodes.

The reference builds ``Outputs_fun(V, P)`` as the horizontal concatenation of ``model.outputs_fun`` over the nodes of
every interval -- shooting node k first (zoh controls), then the d collocation nodes (awebox/ocp/collocation.py:366-381,
awebox/ocp/discretization.py:316-334) -- and reads it by label after every solve (plots, quality checks, the objective's
side-slip term).  Node inputs are the same gathers the constraints use (awebox/tools/struct_operations.py:69-116): the
collocation nodes take ``xdot`` from the collocation polynomial, the shooting nodes take the lifted ``V[xdot, k]``.

``OutputsMap`` holds the expressions (awebox_b200/mdl/outputs.py), their labels and the node order; ``evaluate`` is the
host restatement used by the tests (numpy + expression evaluation, one instance); the product path is the CUDA kernel
behind ``awegpu_eval_outputs`` generated from the same expressions (node_out.inc).
"""
import os

import numpy as np

from .. import sym as S
from ..mdl.outputs import model_outputs


class OutputsMap(object):
    def __init__(self, ocp):
        self.ocp = ocp
        self.entries = model_outputs(ocp.model)              # [(label, Expr)]
        self.labels = [l for l, _ in self.entries]
        self.exprs = [S._e(e) for _, e in self.entries]
        self.n_out = len(self.exprs)
        self.n_nodes = ocp.n_k * (ocp.d + 1)

    def index(self, group, name=None):
        pre = '[%s,' % group if name is None else '[%s,%s,' % (group, name)
        return [i for i, l in enumerate(self.labels) if l.startswith(pre)]

    def node_inputs(self, w, k, j):
        """node variable column [x; xdot; u; z; theta] of node (k, j): j = -1 shooting node, j >= 0 collocation node"""
        ocp = self.ocp
        n_x = ocp.n_x
        if j < 0:
            x = w[ocp.v_x[k]:ocp.v_x[k] + n_x]
            xdot = w[ocp.v_xdot[k]:ocp.v_xdot[k] + n_x]
            z = w[ocp.v_z[k]:ocp.v_z[k] + ocp.n_z]
        else:
            C = ocp.coll.coeff_collocation
            tf = w[ocp.tf_idx[k]]
            Xr = np.array([w[ocp.x_base(k, r):ocp.x_base(k, r) + n_x] for r in range(ocp.d + 1)])
            x = Xr[j + 1]
            xdot = (C[:, j + 1] * ocp.n_k / tf) @ Xr
            z = w[ocp.v_cz[k][j]:ocp.v_cz[k][j] + ocp.n_z]
        u = w[ocp.v_u[k]:ocp.v_u[k] + ocp.n_u]
        theta = w[ocp.theta_idx[k]]
        return np.concatenate([x, xdot, u, z, theta])

    def evaluate(self, w, p):
        """-> [n_nodes, n_out] for one instance (test infrastructure; the product path is the CUDA kernel)"""
        ocp, m = self.ocp, self.ocp.model
        w = np.asarray(w, dtype=float)
        p = np.asarray(p, dtype=float)
        theta0 = p[ocp.p_theta0:] if len(p) == ocp.n_p else p
        params = np.concatenate([theta0, w[ocp.v_phi:ocp.v_phi + ocp.n_phi]])
        leaves = list(m.vars) + list(m.params)
        out = np.zeros((self.n_nodes, self.n_out))
        n = 0
        for k in range(ocp.n_k):
            for j in range(-1, ocp.d):
                v = self.node_inputs(w, k, j)
                out[n] = S.evaluate(self.exprs, leaves, np.concatenate([v, params]))
                n += 1
        return out

    # ------------------------------------------------------------------ code generation
    SECTION_OPS = 0        # operations per section of the straight-line program (0: one section)

    def emit_sections(self, emitter):
        """the program in sections of consecutive labels, each a scope of its own: common subexpressions are shared
        inside a section and recomputed across sections, so that no temporary lives longer than its section (the
        whole program in one scope needs more than 255 registers per thread for the dual-kite model)"""
        budget = int(os.environ.get('AWE_OUT_SECTION_OPS', self.SECTION_OPS))
        if budget <= 0:
            text, st = emitter.emit(self.exprs)
            st = dict(st)
            st['sections'] = 1
            return text, st
        sections, cur = [], []
        for i in range(self.n_out):
            trial = cur + [i]
            if cur and sum(S.count_ops([self.exprs[t] for t in trial]).values()) > budget:
                sections.append(cur)
                cur = [i]
            else:
                cur = trial
        if cur:
            sections.append(cur)
        text, total = '', {}
        for sec in sections:
            body, st = emitter.emit([self.exprs[i] for i in sec])
            for n_local, i in reversed(list(enumerate(sec))):          # local output numbers -> global ones
                body = body.replace('OUT(%d, ' % n_local, 'OUT_(%d, ' % i)
            text += '{\n' + body.replace('OUT_(', 'OUT(') + '}\n'
            for key, v in st.items():
                total[key] = total.get(key, 0) + v
        total['sections'] = len(sections)
        return text, total

    def write(self, out_dir, emitter):
        """node_out.inc (straight-line program: VIN(i) node variables, PIN(i) parameters, OUT(n, value)) and out_defs.h
        (entry count and the label table) for csrc/awegpu_out.cu; ``emitter``: a codegen.NodeCode"""
        text, stats = self.emit_sections(emitter)
        banner = '// generated by awebox_b200/ocp/outputs.py from spec "%s" -- do not edit\n' % self.ocp.spec.name
        head = banner + '#define AWE_N_OUT %d\n' % self.n_out
        head += 'static const char *const awe_out_labels[AWE_N_OUT] = {\n'
        head += ''.join('    "%s",\n' % l for l in self.labels) + '};\n'
        for name, content in (('node_out.inc', banner + text), ('out_defs.h', head)):
            path = os.path.join(out_dir, name)
            if not (os.path.exists(path) and open(path).read() == content):
                with open(path, 'w') as f:
                    f.write(content)
        return stats
