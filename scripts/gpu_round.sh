# round evidence: launch list (B=4096) and one full capture of the dominant kernel (B=4096), each after a plain run
python scripts/profile_step.py 4096 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 4096 3 > gpurun_out/ncu_launch.log 2>&1
cat gpurun_out/plain.log
python scripts/profile_step.py 4096 3 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:awegpu_node_kernel -s 14 -c 1 -o gpurun_out/prof_r1_s0_B4096 python scripts/profile_step.py 4096 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
