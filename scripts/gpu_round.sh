set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
python scripts/profile_step.py 4096 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 4096 3 > gpurun_out/ncu_launch.log 2>&1
cat gpurun_out/plain.log
python scripts/profile_step.py 1024 3 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:awegpu_node_kernel -s 10 -c 5 -o gpurun_out/prof_r1 python scripts/profile_step.py 1024 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
