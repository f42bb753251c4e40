for n in 5 10 20 40 80; do
  echo "ilv_items=$n"; AWEGPU_ILV_ITEMS=$n timeout 120 python scripts/variant_check.py 4096 2>&1 | tail -1 | cut -c41-150
done
