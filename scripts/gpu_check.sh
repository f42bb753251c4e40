for c in 128 256 512 1024 2048; do
  AWEGPU_COPY_CHUNK=$c python bench.py --steps 10 --cpu-seconds 0.5 --e2e-steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('copy chunk $c: e2e %.4e node-evals/s  (%.1f ms/step)' % (d['e2e']['value'], 655360/d['e2e']['value']*1e3))"
done
