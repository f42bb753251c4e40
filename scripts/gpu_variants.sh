for ip in 0 1 2; do
  echo "interleave_pre=$ip"; AWEGPU_INTERLEAVE_PRE=$ip python scripts/variant_check.py 4096 2>&1 | tail -1
done
AWEGPU_INTERLEAVE_PRE=1 python scripts/profile_step.py 4096 3 > /dev/null 2>&1 && AWEGPU_INTERLEAVE_PRE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 18 --csv --log-file gpurun_out/l_cur.csv python scripts/profile_step.py 4096 3 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/l_cur.csv 6
