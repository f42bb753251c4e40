# round evidence with the final binaries: tests, smoke, bench, launch list (B=4096), full capture of the stage-0 kernels
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']['value'], d['e2e_reduced']['value'], d['latency_b1_us']); print({k:v for k,v in d['roofline'].items() if k in ('achieved','frac','traffic','kernel_ms','kernel_share_of_step')}); print(d['cpu_baseline'])"
python scripts/profile_step.py 4096 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 4096 3 > gpurun_out/ncu_launch.log 2>&1
cat gpurun_out/plain.log; python scripts/launch_summary.py gpurun_out/launches.csv | tail -9
python scripts/profile_step.py 4096 3 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"ILb0ELb1ELi0ELb0|ILb1ELb1ELi0ELb0" -s 2 -c 2 -f -o gpurun_out/prof_s0_B4096 python scripts/profile_step.py 4096 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
