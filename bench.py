#!/usr/bin/env python
"""bench.py — collocation-node g + Jacobian evaluations per second (BASELINE.json: metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A step = one pass of the hot path over one batch of synthetic OCP instances.  Workload at N=1:
BASELINE.json configs[1] — the Ampyx AP2 OCP (Radau d=4, n_k=40), batched g + Jacobian at 4096
randomly perturbed primal points.  Weak scaling: every GPU gets its own 4096 instances (batch
sharded, no data-path collective); value = coll-node-evals of all ranks / max-over-ranks time.

``value``  : inputs resident in HBM, outputs left in HBM (device-pointer C-ABI entry).
``e2e``    : the same metric through the host-buffer C-ABI entry (pinned host memory, H2D of W / P
             and D2H of g / jac_g values inside the timed region).
``roofline``: the one kernel of the step, algorithmic bytes / CUDA-event duration vs measured HBM peak.
``cpu_baseline`` / ``--impl reference``: the plain-C checker (oracle/awebox_oracle.c, the CPU port of
             the same path) on the host cores of this box; CasADi itself cannot be installed here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs -> (spec, n_k, instances per GPU (None: total / world), parameter mode, scaling, description)
CONFIGS = {
    'ampyx_ap2': ('ampyx_ap2', 40, 4096, 'shared', 'weak',
                  'ampyx_ap2 OCP (Radau d=4, n_k=40): batched g+jac_g at 4096 perturbed primal points per GPU'),
    'dual_kites': ('dual_kites', 20, 4096, 'shared', 'weak',
                   'dual-kite lift-mode reel-out OCP (Radau d=4, n_k=20): batched g+jac_g at 4096 perturbed primal points per GPU'),
    'power_curve': ('dual_kites', 20, 4096, 'theta0', 'weak',
                    'dual-kite power curve: 64 wind speeds x 64 perturbed primal points per GPU, per-instance theta0 (wind.u_ref)'),
    'mpc_3dof': ('mpc_3dof', 10, None, 'full', 'strong',
                 'tracking MPC (3-DOF, N=10): 8192 scenario evaluations per step over all GPUs, per-instance parameter vectors'),
}
MPC_SCENARIOS = 8192
SPEC, N_K, BATCH = 'ampyx_ap2', 40, 4096
WORKLOAD = CONFIGS['ampyx_ap2'][5]


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_fp64_peak():
    """vector-FP64 peak of this pool's B200 (TFLOP/s), measured with independent DFMA chains on every SM
    (scripts/ubench/ubench.cu, result committed under profiles/)."""
    path = os.path.join(ROOT, 'profiles', 'r02_ubench.jsonl')
    best = None
    if os.path.exists(path):
        for line in open(path):
            try:
                d = json.loads(line)
            except ValueError:
                continue
            if d.get('bench') == 'fp64':
                best = max(best or 0.0, float(d['tflops']))
    return (best, 'measured (profiles/r02_ubench.jsonl)') if best else (37.2, 'nominal (148 SM x 64 FMA/clk x 1.965 GHz)')


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU while the timed region runs (NVML, every few ms;
    nvidia-smi as a fallback when the NVML binding is missing)."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: address the GPU torch uses by its PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id
            dom = torch.cuda.get_device_properties(index).pci_domain_id
            dev = torch.cuda.get_device_properties(index).pci_device_id
            self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(('%08x:%02x:%02x.0' % (dom, bus, dev)).encode())
            self.nvml = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    n = self.nvml
                    sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                            n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
                    self.samples.append([sm, self.sm_max] + ['active' if (r & b) else 'not active' for b in bits])
                    time.sleep(0.002)
                    continue
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                if self.nvml is not None:
                    self.nvml = None        # fall back to nvidia-smi
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(str(s[2 + i]).lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons,
                'samples': len(self.samples), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


def make_inputs(cfg, ocp, B, rank, first=0):
    """synthetic inputs of one rank: W [B, n_v], P (shared vector | [B, n_theta0] | [B, n_p]), cabi parameter flags"""
    from awebox_b200 import synthetic, cabi
    mode = CONFIGS[cfg][3]
    W = synthetic.batch(ocp, B, seed=20261018 + 2 + 1000 * rank)
    if mode == 'shared':
        return W, ocp.default_p(), cabi.AWEGPU_P_SHARED
    if mode == 'theta0':
        return W, synthetic.power_curve_params(ocp, B, first=first), cabi.AWEGPU_P_THETA0
    return W, synthetic.scenario_params(ocp, W, seed=99 + rank), 0


def cpu_leg(ocp, n_threads, budget_s, W, P, flags):
    """time the CPU port on a bounded sample; returns (node-evals/s, seconds, instances, threads)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import build_oracle
    orc = build_oracle.Oracle(ocp)
    n = min(256, W.shape[0])
    sample, psample = W[:n], (P if P.ndim == 1 else P[:n])
    orc.eval(sample, psample, flags=flags, n_threads=n_threads)      # warm-up
    t0 = time.perf_counter()
    done = 0
    while True:
        orc.eval(sample, psample, flags=flags, n_threads=n_threads)
        done += sample.shape[0]
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    return done * ocp.n_k * ocp.d / dt, dt, done, n_threads


def workload_config(name, ocp, B, B_total, p_mode, world, workload, n_k):
    """the `config` object of the bench line: identical for the GPU arm and the reference arm of one command"""
    p_bytes = 8 * (ocp.n_theta0 if p_mode != 'full' else ocp.n_p)
    step_bytes = B * (8 * (ocp.n_v + ocp.n_g + ocp.nnz) + (p_bytes if p_mode != 'shared' else 0))
    return {'workload': workload, 'baseline_config': name, 'n_k': n_k, 'd': ocp.d,
            'instances_per_gpu': B, 'instances_total': B_total, 'n_v': ocp.n_v,
            'n_g': ocp.n_g, 'nnz': int(ocp.nnz), 'jac_layout': 'instance-major CCS values',
            'parameters': {'shared': 'one parameter vector for the batch', 'theta0': 'per-instance theta0 [B, n_theta0]',
                           'full': 'per-instance parameter vectors [B, n_p]'}[p_mode],
            'l2': 'inputs+outputs per step (%.2f GB per GPU) exceed the 126 MB L2; no flush needed' % (step_bytes / 1e9),
            'parallelism': 'batch sharded over %d GPU(s), no data-path collective' % world}


def run_reference(args, json_out):
    """--impl reference: the CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from awebox_b200.spec import ModelSpec
    from awebox_b200.ocp.nlp import OcpProblem
    spec, n_k, per_gpu, p_mode, scaling, workload = CONFIGS[args.config]
    ocp = OcpProblem(ModelSpec.load(spec), n_k=n_k)
    world = max(1, int(args.gpus))
    if per_gpu is None and not args.batch:
        B_cfg, B_total_cfg = -(-MPC_SCENARIOS // world), MPC_SCENARIOS       # strong scaling: rank 0's share
    else:
        B_cfg = args.batch or per_gpu
        B_total_cfg = B_cfg * world
    W, P, flags = make_inputs(args.config, ocp, 256, 0)
    threads = os.cpu_count()
    per_step = max(1.0, min(10.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_leg(ocp, threads, 0.2, W, P, flags)
    secs, inst = 0.0, 0
    for _ in range(args.steps):
        v, dt, n, _ = cpu_leg(ocp, threads, per_step / 2, W, P, flags)
        secs += dt
        inst += n
    value = inst * ocp.n_k * ocp.d / secs
    sample = 'bounded sample: 256-instance blocks of the same workload repeated for %.1f s per step' % (per_step / 2)
    line = {
        'impl': 'reference', 'metric': 'collocation-node g+Jacobian evals/s', 'value': value, 'unit': 'node-evals/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * secs / max(1, args.steps),
        'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.config, ocp, B_cfg, B_total_cfg, p_mode, world, workload, n_k),
        'cpu_baseline': {'value': value, 'unit': 'node-evals/s', 'cores': threads, 'kind': 'port', 'sample': sample,
                         'note': 'plain-C port of the path (gcc -O2, OpenMP over instances); CasADi 3.6.4 itself is '
                                 'not installable on this box'},
        'e2e': {'value': value, 'unit': 'node-evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), file=json_out, flush=True)


def bind_to_gpu_numa_node(index):
    """run this rank (and first-touch its pinned host buffers) on the NUMA node its GPU is attached to: with one
    rank per GPU the host<->device copies of the e2e leg otherwise share one socket's memory controllers.
    Returns (previous affinity, node) or (None, None) when the topology cannot be read."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bdf).read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        before = os.sched_getaffinity(0)
        cpus &= before
        if not cpus:
            return None, None
        os.sched_setaffinity(0, cpus)
        return before, node
    except Exception:
        return None, None


def claim_stdout():
    """stdout carries exactly one JSON line: keep a private handle for it and send everything else that libraries
    write to file descriptor 1 (NCCL's version banner, for instance) to stderr."""
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    return out


def main():
    json_out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='ampyx_ap2', choices=sorted(CONFIGS),
                    help='BASELINE.json configs: ampyx_ap2 = configs[1] (the metric), dual_kites = [2], power_curve = [3], mpc_3dof = [4]')
    ap.add_argument('--batch', type=int, default=0, help='instances per GPU (0: the configuration\'s own)')
    ap.add_argument('--gather', default='none', choices=['none', 'root', 'all'],
                    help='also time the NCCL gather of the per-rank results (g and the reduced products) after a step')
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--layout', default='instance', choices=['instance', 'batch-minor'],
                    help='layout of the jac_g value array on the device (diagnostic; the metric uses instance)')
    ap.add_argument('--no-jac', action='store_true', help='residual only (diagnostic)')
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        return run_reference(args, json_out)

    import numpy as np
    import torch
    import torch.distributed as dist
    from awebox_b200.evaluator import Evaluator
    from awebox_b200 import cabi, parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device visible — this benchmark has no CPU fallback')
    torch.cuda.set_device(local_rank)
    affinity_before, numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else (None, None)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    spec, n_k, per_gpu, p_mode, scaling, workload = CONFIGS[args.config]
    ev = Evaluator(spec, n_k=n_k, device=local_rank)
    ocp = ev.ocp
    if per_gpu is None:
        # strong scaling: the scenario batch is split over the ranks
        lo, hi = parallel.shard_range(MPC_SCENARIOS, rank, world)
        B, B_total, first = hi - lo, MPC_SCENARIOS, lo
    else:
        B = args.batch or per_gpu
        B_total, first = B * world, B * rank
    if args.batch and per_gpu is None:
        B, B_total, first = args.batch, args.batch * world, args.batch * rank
    W, P, pflags = make_inputs(args.config, ocp, B, rank, first=first)
    dW = torch.from_numpy(W).cuda()
    dP = torch.from_numpy(P).cuda()
    dG = torch.empty((B, ocp.n_g), dtype=torch.float64, device='cuda')
    dJ = torch.empty((B, ocp.nnz), dtype=torch.float64, device='cuda')
    stream = torch.cuda.current_stream()

    bm = args.layout == 'batch-minor'

    def step():
        ev.eval_device(dW, dP, dG, None if args.no_jac else dJ, batch_minor=bm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches1 = ev.launch_count()
    kernel_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    gpu_launches = ev.launch_count() - launches1
    # per-launch duration of the one kernel, CUDA events on the launching stream (a few extra launches)
    step_ms = []
    for _ in range(5):
        step()
        torch.cuda.synchronize()
        step_ms.append(ev.last_eval_ms())
        kernel_ms.append(ev.last_kernel_ms())
    sampler.stop_flag = True
    total_ms = max_over_ranks(total_ms)
    nodes_per_step = B_total * ocp.n_k * ocp.d
    value = nodes_per_step * args.steps / (total_ms * 1e-3)

    # ---- device-resident products (J^T lambda, J v, norms): the on-device consumer's view of a step
    rng = np.random.default_rng(4 + rank)
    lam = torch.from_numpy(rng.standard_normal((B, ocp.n_g))).cuda()
    vec = torch.from_numpy(rng.standard_normal((B, ocp.n_v))).cuda()
    red = None
    red_ms = None
    if not args.no_jac and not bm:
        for _ in range(2):
            red = ev.reduce_device(dG, dJ, lam=lam, v=vec)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        for _ in range(5):
            red = ev.reduce_device(dG, dJ, lam=lam, v=vec)
        r1.record(stream)
        torch.cuda.synchronize()
        red_ms = r0.elapsed_time(r1) / 5

    # ---- objective f and grad f (power-cycle configurations): device entry, CUDA events
    obj_line = None
    if p_mode == 'shared' and ocp.traj_type == 'power_cycle':
        from awebox_b200.ocp.objective import Objective
        ob = Objective(ocp)
        p_obj = ocp.default_p()
        p_obj[:ocp.n_v] = W[0]                                        # p.ref: the first primal point
        p_obj[ob.p_weights:ob.p_weights + ocp.n_var] = 1.0
        p_obj[ob.p_cost:ob.p_cost + ob.n_cost] = 1.0
        dPo = torch.from_numpy(p_obj).cuda()
        dF = torch.empty((B,), dtype=torch.float64, device='cuda')
        dGr = torch.empty((B, ocp.n_v), dtype=torch.float64, device='cuda')
        for _ in range(3):
            ev.eval_objective_device(dW, dPo, dF, dGr)
        torch.cuda.synchronize()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(stream)
        for _ in range(10):
            ev.eval_objective_device(dW, dPo, dF, dGr)
        o1.record(stream)
        torch.cuda.synchronize()
        o_ms = o0.elapsed_time(o1) / 10
        o_bytes = B * 8 * (2 * ocp.n_v + 1)                           # read w, write grad f and f
        # the numpy restatement (awebox_b200/ocp/objective.py, the checker of tests/test_objective.py) on the host, one thread
        t0c, n_cpu = time.perf_counter(), 0
        while n_cpu < 4 or (time.perf_counter() - t0c < 1.0 and n_cpu < 64):
            ob.evaluate(W[n_cpu % B], p_obj)
            n_cpu += 1
        cpu_rate = n_cpu / (time.perf_counter() - t0c)
        obj_line = {'ms_per_step': o_ms, 'f_and_grad_evals_per_s': B / (o_ms * 1e-3),
                    'algorithmic_bytes': int(o_bytes), 'GBps': o_bytes / (o_ms * 1e-3) / 1e9,
                    'api': 'awegpu_eval_objective (device pointers): one kernel, CTA = instance',
                    'cpu_port_evals_per_s': cpu_rate, 'cpu_port': 'numpy restatement, 1 thread, %d instances' % n_cpu,
                    'finite': bool(torch.isfinite(dF).all().item())}

    if ocp.traj_type == 'mpc':
        # tracking cost of the periodic MPC (awebox/pmpc.py:304-358): every scenario its own reference and weights
        from awebox_b200.ocp.objective import MpcObjective
        mo = MpcObjective(ocp)
        rng_o = np.random.default_rng(7 + rank)
        Pm = np.empty((B, mo.n_p))
        Pm[:, mo.p_ref:mo.p_ref + ocp.n_v] = W * (1.0 + 1e-2 * rng_o.uniform(-1.0, 1.0, W.shape))
        Pm[:, :mo.p_ref] = 0.0
        Pm[:, mo.p_uref:] = rng_o.uniform(0.5, 2.0, (B, mo.n_p - mo.p_uref))
        dPm = torch.from_numpy(Pm).cuda()
        dF = torch.empty((B,), dtype=torch.float64, device='cuda')
        dGr = torch.empty((B, ocp.n_v), dtype=torch.float64, device='cuda')
        for _ in range(3):
            ev.eval_objective_device(dW, dPm, dF, dGr)
        torch.cuda.synchronize()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(stream)
        for _ in range(10):
            ev.eval_objective_device(dW, dPm, dF, dGr)
        o1.record(stream)
        torch.cuda.synchronize()
        o_ms = o0.elapsed_time(o1) / 10
        o_bytes = B * 8 * (3 * ocp.n_v + 1 + 2 * ocp.n_x + ocp.n_u)   # read w, ref and the weights, write grad f and f
        t0c, n_cpu = time.perf_counter(), 0
        while n_cpu < 4 or (time.perf_counter() - t0c < 0.5 and n_cpu < 2000):
            mo.evaluate(W[n_cpu % B], Pm[n_cpu % B])
            n_cpu += 1
        obj_line = {'ms_per_step': o_ms, 'f_and_grad_evals_per_s': B / (o_ms * 1e-3),
                    'algorithmic_bytes': int(o_bytes), 'GBps': o_bytes / (o_ms * 1e-3) / 1e9,
                    'api': 'awegpu_eval_objective (device pointers), tracking cost of the periodic MPC: one kernel',
                    'cpu_port_evals_per_s': n_cpu / (time.perf_counter() - t0c),
                    'cpu_port': 'numpy restatement, 1 thread, %d instances' % n_cpu,
                    'finite': bool(torch.isfinite(dF).all().item())}

    # ---- second-order hook: triu of the Hessian of lam_f f + lam_g' g, device entry, CUDA events
    hess_line = None
    if p_mode == 'shared':
        Hs = ev.hessian_setup()
        with_obj = ev.hess_objective
        p_h = ocp.default_p()
        if with_obj:
            p_h[:ocp.n_v] = W[0]
            p_h[ocp.n_v:ocp.p_theta0] = 1.0
        dPh = torch.from_numpy(p_h).cuda()
        rng_h = np.random.default_rng(11 + rank)
        dLam = torch.from_numpy(rng_h.uniform(-1.0, 1.0, (B, ocp.n_g))).cuda()
        dSig = torch.ones((B,), dtype=torch.float64, device='cuda') if with_obj else None
        dH = torch.empty((B, Hs.nnz), dtype=torch.float64, device='cuda')
        for _ in range(2):
            ev.eval_hessian_device(dW, dPh, dLam, dH, sigma=dSig)
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        for _ in range(5):
            ev.eval_hessian_device(dW, dPh, dLam, dH, sigma=dSig)
        h1.record(stream)
        torch.cuda.synchronize()
        h_ms = h0.elapsed_time(h1) / 5
        h_bytes = B * 8 * (ocp.n_v + ocp.n_g + Hs.nnz)               # read w and lam, write the non-zeros of triu(H)
        # the numpy restatement (awebox_b200/ocp/hessian.py, the checker of tests/test_hessian.py) on the host, one thread
        t0c, n_cpu = time.perf_counter(), 0
        while n_cpu < 2 or (time.perf_counter() - t0c < 1.0 and n_cpu < 32):
            Hs.evaluate(W[n_cpu % B], p_h, dLam[n_cpu % B].cpu().numpy(), 1.0 if with_obj else 0.0)
            n_cpu += 1
        hess_line = {'ms_per_step': h_ms, 'hess_lag_evals_per_s': B / (h_ms * 1e-3),
                     'node_hessians_per_s': B * ocp.n_k * (ocp.d + 1) / (h_ms * 1e-3), 'nnz_triu': int(Hs.nnz),
                     'with_objective_part': bool(with_obj),
                     'algorithmic_bytes': int(h_bytes), 'GBps': h_bytes / (h_ms * 1e-3) / 1e9,
                     'api': 'awegpu_eval_hessian (device pointers): node Hessian kernel + assembly kernel per chunk of 2048 instances',
                     'cpu_port_evals_per_s': n_cpu / (time.perf_counter() - t0c),
                     'cpu_port': 'numpy restatement, 1 thread, %d instances' % n_cpu,
                     'finite': bool(torch.isfinite(dH).all().item())}
        del dH, dLam

    # ---- outputs map: nlp.Outputs_fun over all nodes, device entry, CUDA events
    out_line = None
    if p_mode == 'shared':
        n_out, n_nodes = len(ev.outputs_labels()), ocp.n_k * (ocp.d + 1)
        dPo_ = torch.from_numpy(ocp.default_p()).cuda()
        dOut = torch.empty((B, n_out, n_nodes), dtype=torch.float64, device='cuda')
        for _ in range(2):
            ev.eval_outputs_device(dW, dPo_, dOut)
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record(stream)
        for _ in range(5):
            ev.eval_outputs_device(dW, dPo_, dOut)
        q1.record(stream)
        torch.cuda.synchronize()
        q_ms = q0.elapsed_time(q1) / 5
        q_bytes = B * 8 * (ocp.n_v + n_out * n_nodes)                 # read w, write every output of every node
        # the host restatement (awebox_b200/ocp/outputs.py, the checker of tests/test_outputs.py), one thread
        from awebox_b200.ocp.outputs import OutputsMap
        om = OutputsMap(ocp)
        t0c, n_cpu = time.perf_counter(), 0
        while n_cpu < 1 or (time.perf_counter() - t0c < 1.0 and n_cpu < 32):
            om.evaluate(W[n_cpu % B], ocp.default_p())
            n_cpu += 1
        out_line = {'ms_per_step': q_ms, 'outputs_fun_evals_per_s': B / (q_ms * 1e-3),
                    'node_outputs_per_s': B * n_nodes / (q_ms * 1e-3), 'n_out': n_out, 'n_nodes': n_nodes,
                    'algorithmic_bytes': int(q_bytes), 'GBps': q_bytes / (q_ms * 1e-3) / 1e9,
                    'api': 'awegpu_eval_outputs (device pointers): one kernel, thread = node, Out[b][label][node]',
                    'cpu_port_evals_per_s': n_cpu / (time.perf_counter() - t0c),
                    'cpu_port': 'expression-graph restatement, 1 thread, %d instances' % n_cpu,
                    'finite': bool(torch.isfinite(dOut).all().item())}
        del dOut

    # ---- NCCL gather of the per-rank results (north_star: "NCCL over NVLink used only to gather per-shard results")
    gather = None
    if args.gather != 'none' and world > 1:
        def do_gather(t):
            return parallel.gather_to_root(t, B_total) if args.gather == 'root' else parallel.gather_results(t, B_total)
        items = {'g': dG}
        if red is not None:
            items.update({'JTl': red['JTl'], 'Jv': red['Jv']})
        gather = {'mode': args.gather}
        for key, t in items.items():
            do_gather(t)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            for _ in range(3):
                out = do_gather(t)
            g1.record(stream)
            barrier()
            ms = max_over_ranks(g0.elapsed_time(g1) / 3)
            nbytes = B_total * t.shape[1] * 8
            gather[key] = {'ms': ms, 'bytes': int(nbytes), 'GBps_into_a_rank': nbytes * (world - 1) / world / (ms * 1e-3) / 1e9}
            del out
        gather['limiter'] = ('the receiving GPU\'s NVLink ingress (nominal 900 GB/s per direction, 770 GB/s measured peer copy)'
                             ' and NCCL launch latency at these sizes')
        # the full Jacobian value arrays are not gathered: they stay sharded for the on-device consumer

    # ---- end to end through the host-buffer entry (pinned host memory)
    hW = torch.from_numpy(W).pin_memory()
    hP = torch.from_numpy(P).pin_memory()
    hG = torch.empty((B, ocp.n_g), dtype=torch.float64).pin_memory()
    hJ = torch.empty((B, ocp.nnz), dtype=torch.float64).pin_memory()
    flags = pflags | cabi.AWEGPU_HOST_PINNED

    def e2e_step():
        rc = ev.lib.awegpu_eval_batch_host(ev.handle, B, hW.data_ptr(), hP.data_ptr(), hG.data_ptr(), hJ.data_ptr(), flags)
        if rc != 0:
            raise SystemExit('awegpu_eval_batch_host failed: %s' % ev.lib.awegpu_last_error(ev.handle).decode())

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_value = nodes_per_step * args.e2e_steps / max_over_ranks(time.perf_counter() - t0)
    assert bool(np.isfinite(hG.numpy()).all())
    del hJ

    # ---- the same with the reduced result set: g, J^T lambda and max|g| cross PCIe (what a host-side outer loop
    #      monitors: primal and dual residuals), the Jacobian values stay on the device for the batched consumer
    hL = torch.from_numpy(rng.standard_normal(ocp.n_g)).pin_memory()
    hJTl = torch.empty((B, ocp.n_v), dtype=torch.float64).pin_memory()
    hGi = torch.empty((B,), dtype=torch.float64).pin_memory()

    def e2e_red_step():
        rc = ev.lib.awegpu_eval_batch_reduce_host(ev.handle, B, hW.data_ptr(), hP.data_ptr(), hL.data_ptr(), None,
                                                  flags, cabi.AWEGPU_RED_LAMBDA_SHARED, hG.data_ptr(), hJTl.data_ptr(), None,
                                                  None, None, hGi.data_ptr())
        if rc != 0:
            raise SystemExit('awegpu_eval_batch_reduce_host failed: %s' % ev.lib.awegpu_last_error(ev.handle).decode())

    e2e_red_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_red_step()
    barrier()
    e2e_red_value = nodes_per_step * args.e2e_steps / max_over_ranks(time.perf_counter() - t0)
    assert bool(np.isfinite(hJTl.numpy()).all())

    # ---- one instance per call (what one IPOPT callback costs): g alone, g + jac_g, host numpy in and out
    lat = {}
    if rank == 0:
        w1 = W[0].copy()
        p1 = P if P.ndim == 1 else P[0].copy()
        for name, fn in (('g_fun', lambda: ev.g_fun(w1, p1)), ('jac_g_fun', lambda: ev.jac_g_fun(w1, p1))):
            for _ in range(20):
                fn()
            t0 = time.perf_counter()
            for _ in range(200):
                fn()
            lat[name] = (time.perf_counter() - t0) / 200 * 1e6
        if p_mode == 'shared':
            # the other functions of one IPOPT iteration / one post-processing pass, one instance per call
            lam1 = np.random.default_rng(3).uniform(-1.0, 1.0, ocp.n_g)
            more = [('outputs_fun', lambda: ev.outputs_fun(w1, ocp.default_p()))]
            ev.hessian_setup()
            p_h1 = ocp.default_p()
            if ev.hess_objective:
                p_h1[:ocp.n_v] = w1
                p_h1[ocp.n_v:ocp.p_theta0] = 1.0
                more.append(('f_and_grad_f', lambda: ev.eval_objective(w1[None, :], p_h1)))
            more.append(('hess_lag_fun', lambda: ev.hess_lag_fun(w1, p_h1, 1.0, lam1)))
            for name, fn in more:
                for _ in range(5):
                    fn()
                t0 = time.perf_counter()
                for _ in range(50):
                    fn()
                lat[name] = (time.perf_counter() - t0) / 50 * 1e6

    if rank == 0:
        peak, peak_src = measured_peak()
        fp64_peak, fp64_src = measured_fp64_peak()
        # dominant kernel: the output stage of the collocation-node programs (awegpu_node_kernel<collocation, full,
        # stage 0>), one launch per step.  Algorithmic bytes of that launch: every collocation node reads
        # its input column once ([x; xdot; z] per node, [u; theta; phi] per interval) and writes its
        # residual rows and Jacobian entries once.
        k_ms = float(np.median(kernel_ms))
        s_ms = float(np.median(step_ms))
        entries = int((ocp.coll_slot >= 0).sum())                       # Jacobian entries of all coll nodes, one instance
        per_inst = 8 * (entries + ocp.n_k * ocp.d * (ocp.n_eq + 2 * ocp.n_x + ocp.n_z)
                        + ocp.n_k * (ocp.n_u + ocp.n_theta + ocp.n_phi))
        alg_bytes = B * per_inst
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        info = ev.kernel_info()
        flops = B * (ocp.n_k * ocp.d * ev.flops_coll + ocp.n_k * ev.flops_shoot)
        executed = B * (ocp.n_k * ocp.d * info['ops_coll_node'] + ocp.n_k * info['ops_shoot_node'])
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath) and args.config == 'ampyx_ap2':
            with open(tpath) as f:
                traffic = json.load(f).get('dram_bytes_per_launch')
        p_bytes = 8 * (ocp.n_theta0 if p_mode != 'full' else ocp.n_p)
        step_bytes = B * (8 * (ocp.n_v + ocp.n_g + ocp.nnz) + (p_bytes if p_mode != 'shared' else 0))
        if affinity_before is not None:
            os.sched_setaffinity(0, affinity_before)          # the CPU leg uses every host core again
        cpu_v, cpu_dt, cpu_n, cpu_thr = cpu_leg(ocp, os.cpu_count(), args.cpu_seconds, W, P, pflags)
        t_one = time.perf_counter()
        cpu_leg(ocp, 1, 0.0, W[:1], P if P.ndim == 1 else P[:1], pflags)
        red_bytes = 8 * (ocp.n_v + ocp.n_g + 1)
        line = {
            'metric': 'collocation-node g+Jacobian evals/s', 'value': value, 'unit': 'node-evals/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.config, ocp, B, B_total, p_mode, world, workload, n_k),
            'kernels_per_step': 'gather + %d collocation stages + %d shooting stages' % (
                sum(1 for g in info['groups_coll'] if g), sum(1 for g in info['groups_shoot'] if g)),
            'clocks': sampler.summary(),
            'e2e': {'value': e2e_value, 'unit': 'node-evals/s',
                    'h2d_bytes_per_step': int(8 * (W.size + P.size)) * world,
                    'd2h_bytes_per_step': int(8 * B * (ocp.n_g + ocp.nnz)) * world,
                    'api': 'awegpu_eval_batch_host (pinned host buffers)', 'timing': 'wall clock around the synchronous calls, max over ranks',
                    'numa_node_of_rank0': numa_node},
            'e2e_reduced': {'value': e2e_red_value, 'unit': 'node-evals/s',
                            'h2d_bytes_per_step': int(8 * (W.size + P.size + ocp.n_g)) * world,
                            'd2h_bytes_per_step': int(B * red_bytes) * world,
                            'api': 'awegpu_eval_batch_reduce_host (pinned host buffers, chunked over three streams): g, J^T lambda '
                                   'and max|g| per instance cross PCIe; the jac_g values stay on the device for a batched consumer',
                            'reduce_kernel_ms': red_ms,
                            'reduce_kernel': 'awegpu_reduce_kernel, all five products (J^T lambda, J v, row / column max-norms, '
                                             'max|g|) with per-instance lambda and v: one pass over the jac_g values'},
            'latency_b1_us': {'g_fun': lat.get('g_fun'), 'jac_g_fun': lat.get('jac_g_fun'),
                              'f_and_grad_f': lat.get('f_and_grad_f'), 'hess_lag_fun': lat.get('hess_lag_fun'),
                              'outputs_fun': lat.get('outputs_fun'),
                              'what': 'one OCP instance per call through Evaluator.g_fun / jac_g_fun (host numpy in and out), '
                                      'mean of 200 calls; the CPU port needs cpu_port_us on one core',
                              'cpu_port_us': (time.perf_counter() - t_one) * 1e6 / 2},
            'gpu_launches': int(gpu_launches),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'kernel': 'awegpu_node_kernel<collocation, full, stage 0> (output stage of the collocation-node programs)',
                         'kernel_ms': k_ms, 'kernel_share_of_step': k_ms / s_ms,
                         'algorithmic_bytes_per_launch': int(alg_bytes),
                         'algorithmic_bytes_per_node': per_inst / (ocp.n_k * ocp.d),
                         'step': {'ms': s_ms, 'algorithmic_bytes': int(step_bytes),
                                  'achieved': step_bytes / (s_ms * 1e-3) / 1e9, 'frac': step_bytes / (s_ms * 1e-3) / 1e9 / peak},
                         'fp64': {'algorithmic_flops_per_step': int(flops), 'executed_flops_per_step': int(executed),
                                  'executed_over_algorithmic': executed / max(1, flops),
                                  'tflops_executed': executed / (s_ms * 1e-3) / 1e12,
                                  'peak_tflops': fp64_peak, 'frac': executed / (s_ms * 1e-3) / 1e12 / fp64_peak,
                                  'peak_source': fp64_src},
                         'programs': info},
            'cpu_baseline': {'value': cpu_v, 'unit': 'node-evals/s', 'cores': cpu_thr, 'kind': 'port',
                             'sample': '256-instance blocks of the same workload repeated for %.1f s (%d instances)'
                                       % (cpu_dt, cpu_n)},
        }
        if obj_line is not None:
            obj_line['frac_of_hbm_peak'] = obj_line['GBps'] / peak
            line['objective'] = obj_line
        if hess_line is not None:
            hess_line['frac_of_hbm_peak'] = hess_line['GBps'] / peak
            line['hessian'] = hess_line
        if out_line is not None:
            out_line['frac_of_hbm_peak'] = out_line['GBps'] / peak
            line['outputs'] = out_line
        if gather is not None:
            line['gather'] = gather
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
