python scripts/profile_step.py 4096 3 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:awegpu_coll_ws_kernel -s 2 -c 1 -o gpurun_out/prof_r1_s0_B4096 python scripts/profile_step.py 4096 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
