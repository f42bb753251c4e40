"""Build reference (awebox) trials through the CasADi restatement (TEST INFRASTRUCTURE — oracle).

Each ``config_*`` returns the option seed of one BASELINE.json configuration, following the
reference's own example scripts; ``build`` runs the reference's Options / Architecture / Model /
Formulation / NLP builders (awebox/trial.py:114-138) and the reference's own initial guess
(awebox/opti/preparation.py:45-79).  Only usable where /root/reference exists.
"""
import numpy as np
import ref_env


def config_ampyx_ap2(n_k=40):
    """examples/ampyx_ap2_trajectory.py:25-48 (BASELINE.json configs[0], configs[1])."""
    import awebox.opts.kite_data.ampyx_ap2_settings as ampyx_ap2_settings
    options = {}
    options['user_options.system_model.architecture'] = {1: 0}
    options = ampyx_ap2_settings.set_ampyx_ap2_settings(options)
    options['user_options.trajectory.type'] = 'power_cycle'
    options['user_options.trajectory.system_type'] = 'lift_mode'
    options['user_options.trajectory.lift_mode.windings'] = 1
    options['params.wind.z_ref'] = 100.0
    options['params.wind.power_wind.exp_ref'] = 0.15
    options['user_options.wind.model'] = 'power'
    options['user_options.wind.u_ref'] = 10.
    options['nlp.n_k'] = n_k
    options['nlp.collocation.u_param'] = 'zoh'
    options['user_options.trajectory.lift_mode.phase_fix'] = 'simple'
    options['solver.linear_solver'] = 'ma57'
    return options


def config_dual_kites(n_k=20):
    """examples/dual_kites_power_curve.py:22-42 (BASELINE.json configs[2], configs[3])."""
    import awebox.opts.kite_data.ampyx_ap2_settings as ampyx_ap2_settings
    options = {}
    options['user_options.system_model.architecture'] = {1: 0, 2: 1, 3: 1}
    options = ampyx_ap2_settings.set_ampyx_ap2_settings(options)
    options['user_options.trajectory.type'] = 'power_cycle'
    options['user_options.trajectory.system_type'] = 'lift_mode'
    options['user_options.trajectory.lift_mode.windings'] = 1
    options['model.system_bounds.theta.t_f'] = [5., 15.]
    options['params.wind.z_ref'] = 100.0
    options['params.wind.power_wind.exp_ref'] = 0.15
    options['user_options.wind.model'] = 'power'
    options['user_options.wind.u_ref'] = 10.
    options['nlp.n_k'] = n_k
    options['nlp.collocation.u_param'] = 'zoh'
    options['user_options.trajectory.lift_mode.phase_fix'] = 'single_reelout'
    options['solver.linear_solver'] = 'ma57'
    return options


def config_mpc_3dof(n_k=10):
    """examples/mpc_closed_loop.py:19-70 through awebox/pmpc.py:68-84 (BASELINE.json configs[4]): the tracking
    MPC trial Pmpc builds from the periodic trial's option seed — single 3-DOF kite, trajectory type 'mpc',
    horizon N = 10, Radau d = 4, zoh controls, theta fixed (t_f = N * ts, ts = 0.1 s).  Pmpc takes the other
    theta entries from the solved periodic trial; here they are the reference's own initial-guess values
    (diam_t 5e-3), since IPOPT is not available to solve the periodic trial."""
    import awebox.opts.kite_data.ampyx_data as ampyx_data
    options = {}
    options['user_options.system_model.architecture'] = {1: 0}
    options['user_options.kite_standard'] = ampyx_data.data_dict()
    options['user_options.system_model.kite_dof'] = 3
    options['user_options.trajectory.system_type'] = 'lift_mode'
    options['user_options.trajectory.lift_mode.windings'] = 1
    options['model.system_bounds.theta.t_f'] = [5.0, 15.0]
    options['params.wind.z_ref'] = 10.0
    options['user_options.wind.model'] = 'log_wind'
    options['user_options.wind.u_ref'] = 5.
    options['user_options.trajectory.lift_mode.phase_fix'] = 'simple'
    options['solver.linear_solver'] = 'ma57'
    options['solver.mu_hippo'] = 1e-2
    # awebox/pmpc.py:68-84
    options['user_options.trajectory.type'] = 'mpc'
    options['nlp.discretization'] = 'direct_collocation'
    options['nlp.n_k'] = n_k
    options['nlp.collocation.d'] = 4
    options['nlp.collocation.scheme'] = 'radau'
    options['nlp.collocation.u_param'] = 'zoh'
    options['formulation.mpc.terminal_point_constr'] = False
    options['user_options.trajectory.fixed_params'] = {'diam_t': 5e-3, 't_f': n_k * 0.1}
    return options


CONFIGS = {'ampyx_ap2': config_ampyx_ap2, 'dual_kites': config_dual_kites, 'mpc_3dof': config_mpc_3dof}


class RefTrial(object):
    pass


def build(name, n_k=None, with_nlp=True, overwrite=None):
    ref_env.activate()
    import awebox.opts.options as opts
    import awebox.mdl.architecture as archi
    import awebox.mdl.model as model
    import awebox.ocp.formulation as formulation
    import awebox.ocp.nlp as nlp
    seed = CONFIGS[name](n_k) if n_k is not None else CONFIGS[name]()
    for k, v in (overwrite or {}).items():
        seed[k] = v
    t = RefTrial()
    t.options = opts.Options()
    t.options.fill_in_seed(seed)
    t.architecture = archi.Architecture(t.options['user_options']['system_model']['architecture'])
    t.options.build(t.architecture)
    t.model = model.Model()
    t.model.build(t.options['model'], t.architecture)
    if with_nlp:
        t.formulation = formulation.Formulation()
        t.formulation.build(t.options['formulation'], t.model)
        t.nlp = nlp.NLP()
        t.nlp.build(t.options['nlp'], t.model, t.formulation)
        import awebox.opti.preparation as prep
        import awebox.opti.initialization_dir.initialization as initialization
        t.p_fix_num = prep.initialize_opti_parameters_with_model_parameters(t.nlp, t.options['solver'])
        t.V_init = initialization.get_initial_guess(t.nlp, t.model, t.formulation,
                                                    t.options['solver']['initialization'], t.p_fix_num)
        t.w0 = np.array(t.V_init.cat.full()).ravel()
        t.p0 = np.array(t.p_fix_num.cat.full()).ravel()
    return t
