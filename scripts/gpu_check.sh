python scripts/variant_check.py 4096 ampyx_ap2 40 2>&1 | tail -1
python scripts/variant_check.py 257 ampyx_ap2 4 2>&1 | tail -1
python -m pytest tests -x -q -m gpu -k "ampyx or ragged or empty or layouts or reference_style or full_size or native" 2>&1 | tail -4
python scripts/profile_step.py 4096 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 18 --csv --log-file gpurun_out/l_cur.csv python scripts/profile_step.py 4096 3 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/l_cur.csv 6
