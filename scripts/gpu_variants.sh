for f in awebox_b200/lib/var/*.so; do
  AWEGPU_LIB=$f timeout 300 python scripts/variant_check.py 4096 2>&1 | tail -1 | cut -c1-40,41-150
done
