python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1
python scripts/variant_check.py 4096 dual_kites 20 2>&1 | tail -1
python scripts/variant_check.py 8192 mpc_3dof 10 2>&1 | tail -1
