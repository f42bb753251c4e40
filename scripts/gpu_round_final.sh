# final round-2 evidence: smoke, bench (both arms), launch list of the bench command, full captures of the outputs
# kernel and of the Hessian kernels (each only after its command has exited 0 without ncu)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']['value'], d['e2e_reduced']['value'], d['latency_b1_us']); print({k:v for k,v in d['roofline'].items() if k in ('achieved','frac','traffic','kernel_ms','kernel_share_of_step')}); print(d['cpu_baseline']); print(d.get('outputs')); print(d.get('hessian')); print(d.get('objective'))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; head -c 600 gpurun_out/bench_ref.json
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_bench.csv | tail -16
python scripts/out_time.py --reps 2 > gpurun_out/plain_out.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:awegpu_out_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_out python scripts/out_time.py --reps 2 > gpurun_out/ncu_out.log 2>&1
cat gpurun_out/plain_out.log; tail -2 gpurun_out/ncu_out.log
python scripts/hess_time.py --reps 2 > gpurun_out/plain_hess.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:awegpu_hess -s 4 -c 2 -f -o gpurun_out/r02_prof_hess python scripts/hess_time.py --reps 2 > gpurun_out/ncu_hess2.log 2>&1
cat gpurun_out/plain_hess.log; tail -2 gpurun_out/ncu_hess2.log
