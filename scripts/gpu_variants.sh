for f in awebox_b200/lib/var/*.so; do
  AWEGPU_LIB=$f timeout 300 python scripts/variant_check.py 4096 2>&1 | tail -1
  AWEGPU_LIB=$f python scripts/profile_step.py 4096 3 > /dev/null 2>&1 && AWEGPU_LIB=$f ncu --metrics gpu__time_duration.sum --clock-control none -c 18 --csv --log-file gpurun_out/l_$(basename $f).csv python scripts/profile_step.py 4096 3 > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/l_$(basename $f).csv 6
done
