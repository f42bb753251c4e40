# usage: gpu_ncu_variant.sh <lib> <tag> [batch]
LIB=$1; TAG=$2; B=${3:-1024}
AWEGPU_LIB=$LIB python scripts/profile_step.py $B 3 > gpurun_out/plain_$TAG.log 2>&1 && \
AWEGPU_LIB=$LIB ncu --set full --clock-control none --import-source on -k regex:awegpu_node_kernel -s 14 -c 1 -o gpurun_out/prof_$TAG python scripts/profile_step.py $B 3 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log; cat gpurun_out/plain_$TAG.log
