timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 120 python scripts/variant_check.py 4096 ampyx_ap2 40 2>&1 | tail -1
AWEGPU_WS=0 timeout 120 python scripts/variant_check.py 4096 ampyx_ap2 40 2>&1 | tail -1
timeout 120 python scripts/variant_check.py 4096 dual_kites 20 2>&1 | tail -1
AWEGPU_WS=0 timeout 120 python scripts/variant_check.py 4096 dual_kites 20 2>&1 | tail -1
timeout 120 python scripts/variant_check.py 8192 mpc_3dof 10 2>&1 | tail -1
AWEGPU_WS=0 timeout 120 python scripts/variant_check.py 8192 mpc_3dof 10 2>&1 | tail -1
