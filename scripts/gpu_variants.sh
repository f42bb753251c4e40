for f in awebox_b200/lib/var/*.so; do
  AWEGPU_LIB=$f timeout 90 python scripts/variant_check.py 4096 2>&1 | tail -2 | cut -c1-150
  echo "rc=$?"
done
nvidia-smi --query-gpu=name,utilization.gpu --format=csv,noheader
