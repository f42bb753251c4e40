timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
