# time the Hessian and the outputs map with every variant library under awebox_b200/lib/var
for f in awebox_b200/lib/var/libawegpu_${AWE_SPEC:-ampyx_ap2}.*.so; do
  echo "== $f"
  AWEGPU_LIB=$f timeout 200 python scripts/hess_time.py --config ${AWE_SPEC:-ampyx_ap2} 2>&1 | tail -1
  AWEGPU_LIB=$f timeout 200 python scripts/out_time.py --config ${AWE_SPEC:-ampyx_ap2} 2>&1 | tail -1
done
