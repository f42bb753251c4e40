python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print({k:v for k,v in d['roofline'].items() if k in ('achieved','frac','traffic','kernel_ms','kernel_share_of_step')}); print(d['cpu_baseline'])"
python scripts/variant_check.py 4096 dual_kites 20 2>&1 | tail -1
python scripts/variant_check.py 8192 mpc_3dof 10 2>&1 | tail -1
