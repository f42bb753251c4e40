# usage: gpu_ncu_variant.sh <lib> <tag> <kernel regex> <skip> [batch]
LIB=$1; TAG=$2; KRE=$3; SKIP=$4; B=${5:-1024}
AWEGPU_LIB=$LIB python scripts/profile_step.py $B 3 > gpurun_out/plain_$TAG.log 2>&1 && \
AWEGPU_LIB=$LIB ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o gpurun_out/prof_$TAG python scripts/profile_step.py $B 3 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log; cat gpurun_out/plain_$TAG.log
