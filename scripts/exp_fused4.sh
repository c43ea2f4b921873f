run() {
  echo -n "$*: "
  env "$@" timeout 120 python bench.py --frames 4096 --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0 $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:v['mean_ms'] for k,v in d['kernels'].items()})"
}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run UPSP_FUSED_V=3
run UPSP_FUSED_V=4
run UPSP_FUSED_BS=64
run UPSP_FUSED_BS=256
run UPSP_FUSED_BS=32
run UPSP_FUSED_BS=128 UPSP_FUSED_OCC=10
