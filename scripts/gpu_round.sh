# round evidence with the final binaries: tests, smoke, bench, launch list (B=4096), full capture of the dominant kernel
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print({k:v for k,v in d['roofline'].items() if k in ('achieved','frac','traffic','kernel_ms','kernel_share_of_step')}); print(d['cpu_baseline'])"
python scripts/profile_step.py 4096 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 4096 3 > gpurun_out/ncu_launch.log 2>&1
cat gpurun_out/plain.log
python scripts/profile_step.py 4096 3 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:awegpu_coll_ws_kernel -s 2 -c 1 -o gpurun_out/prof_r1_s0_B4096 python scripts/profile_step.py 4096 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python scripts/variant_check.py 4096 dual_kites 20 2>&1 | tail -1
python scripts/variant_check.py 8192 mpc_3dof 10 2>&1 | tail -1
