#!/bin/bash
# time + parity of every variant library under awebox_b200/lib/var (scripts/build_variants.py)
for f in awebox_b200/lib/var/libawegpu_${AWE_SPEC:-ampyx_ap2}.*.so; do
  AWEGPU_LIB=$f timeout 120 python scripts/variant_check.py ${1:-4096} ${AWE_SPEC:-ampyx_ap2} 2>&1 | tail -1
done
