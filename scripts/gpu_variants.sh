for f in awebox_b200/lib/var/*.so; do
  AWEGPU_LIB=$f timeout 300 python scripts/variant_check.py 4096 2>&1 | tail -1
done
for ct in 4 8 16 32; do
  echo "chunk_tiles=$ct"; AWEGPU_CHUNK_TILES=$ct AWEGPU_LIB=awebox_b200/lib/var/libawegpu_ampyx_ap2.nopf.so timeout 300 python scripts/variant_check.py 4096 2>&1 | tail -1
done
for il in 0 2 4; do
  echo "interleave=$il"; AWEGPU_INTERLEAVE=$il AWEGPU_LIB=awebox_b200/lib/var/libawegpu_ampyx_ap2.nopf.so timeout 300 python scripts/variant_check.py 4096 2>&1 | tail -1
done
