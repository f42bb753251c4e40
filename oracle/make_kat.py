"""Known-answer fixtures of the reference's own unit tests for the node dynamics.  TEST INFRASTRUCTURE.

    PYTHONHASHSEED=0 python oracle/make_kat.py

/root/reference/test/units/test_model.py:473-772 checks ``Model.dynamics`` on two systems with a closed-form
solution — a pendulum and a "pseudo-Atwood" machine (a mass on a tether that is reeled out), with a massless
and with a heavy tether — by building a 3-DOF single-kite model without aerodynamic forces and evaluating the
residual at analytically consistent inputs.  This script imports that test module from /root/reference over the
CasADi restatement, runs its builders unchanged and stores, for every case,

  tests/golden/kat_<case>.json   the numeric ModelSpec of the test model, the input vectors the reference test
                                 populated (variables, parameters), the residual the reference computed and the
                                 tolerance the reference asserts.

tests/test_reference_kats.py evaluates the product's node model on the same inputs.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_env                                    # noqa: E402

CASES = [
    # name, system, consistent inputs, heavy tether, tolerance asserted by the reference test (test_model.py:669-730)
    ('pendulum_massless', 'pendulum', True, False, 1e-5),
    ('pendulum_heavy_rod', 'pendulum', True, True, 1e-5),
    ('pendulum_inconsistent', 'pendulum', False, False, 1e-2),
    ('atwood_massless', 'pseudo_atwood', True, False, 1e-5),
    ('atwood_heavy_tether', 'pseudo_atwood', True, True, 1e-2),
    ('atwood_inconsistent', 'pseudo_atwood', False, False, 1e-2),
]


def main():
    import logging
    logging.disable(logging.CRITICAL)
    cas = ref_env.activate()
    sys.path.insert(0, os.path.join(ref_env.REFERENCE, 'test', 'units'))
    import test_model as ref_test
    import awebox.tools.vector_operations as vect_op
    import awebox as awe
    import awebox.mdl.architecture as archi
    from awebox_b200 import spec as bspec
    for name, system, consistent, heavy, eps in CASES:
        pars = ref_test.get_arbitary_system_parameters(rod_has_mass=heavy, pendulum_or_pseudo_atwood=system)
        if system == 'pendulum':
            initial_si = ref_test.get_consistent_inputs_for_pendulum_problem(pars)
        else:
            initial_si = ref_test.get_consistent_inputs_for_pseudo_atwood_problem(pars)
        if not consistent:                          # test_model.py:745-748
            initial_si['q10'] = 80 * vect_op.xhat_dm()
            initial_si['dq10'] = 50. + initial_si['q10']
            initial_si['lambda10'] = 4.
        # build_pendulum_or_pseudo_atwood_test_model (test_model.py:573-598), keeping the option tree for the spec
        options = {}
        options['user_options.system_model.architecture'] = {1: 0}
        options['user_options.system_model.kite_dof'] = 3
        options['user_options.wind.model'] = 'uniform'
        options['user_options.atmosphere'] = 'uniform'
        options['user_options.induction_model'] = 'not_in_use'
        options['user_options.tether_drag_model'] = 'not_in_use'
        options['user_options.kite_standard'] = awe.ampyx_data.data_dict()
        options['model.geometry.overwrite.m_k'] = pars['mass_si']
        options['model.tether.control_var'] = 'ddl_t'
        options['params.atmosphere.g'] = pars['gravity_si']
        options['params.tether.rho'] = pars['rod_density_si']
        trial_options = awe.Options()
        trial_options.fill_in_seed(options)
        architecture = archi.Architecture(trial_options['user_options']['system_model']['architecture'])
        trial_options.build(architecture)
        model = awe.mdl.model.Model()
        model.build(trial_options['model'], architecture)
        variables, parameters = ref_test.populate_model_variables_and_parameters(model, pars, initial_si)
        res = np.array(model.dynamics(variables, parameters).full()).ravel()
        v = np.array(variables.cat.full()).ravel()
        p = np.array(parameters.cat.full()).ravel()
        sp = bspec.from_awebox('kat_' + name, trial_options, architecture, model,
                               theta0_values=[float(x) for x in p[:len([l for l in model.parameters.labels() if l.startswith('[theta0,')])]])
        out = {'spec': sp.d, 'variables': v.tolist(), 'parameters': p.tolist(), 'residual': res.tolist(),
               'eq_names': list(model.constraints_list.get_name_list('eq')), 'epsilon': eps, 'consistent': consistent,
               'reference_test': 'test/units/test_model.py: check_dynamics_with_known_inputs_in_known_situations('
                                 'use_consistent_inputs=%s, rod_has_mass=%s, pendulum_or_pseudo_atwood=%r)' % (consistent, heavy, system)}
        path = os.path.join(ROOT, 'tests', 'golden', 'kat_%s.json' % name)
        with open(path, 'w') as f:
            json.dump(out, f, indent=1, sort_keys=True)
        print('%-24s |residual| = %.3e (eps %.0e, consistent=%s) -> %s' % (name, np.linalg.norm(res), eps, consistent,
                                                                           os.path.relpath(path, ROOT)))


if __name__ == '__main__':
    main()
