for d in 0 1 2 4 3 7; do
  echo -n "dbg=$d: "
  UPSP_FUSED_DBG=$d timeout 120 python bench.py --frames 4096 --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:v['mean_ms'] for k,v in d['kernels'].items()})"
done
