for rep in 1 2 3; do
for f in awebox_b200/lib/var/*pol0.so awebox_b200/lib/var/*pol1.so; do
  AWEGPU_LIB=$f timeout 300 python scripts/variant_check.py 4096 2>&1 | tail -1 | cut -c1-40,86-140
done
done
