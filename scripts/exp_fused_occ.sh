run() {
  echo -n "$*: "
  env "$@" timeout 120 python bench.py --frames 4096 --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0 $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:v['mean_ms'] for k,v in d['kernels'].items()})"
}
run UPSP_FUSED_BS=128 UPSP_FUSED_OCC=8
run UPSP_FUSED_BS=128 UPSP_FUSED_OCC=10
run UPSP_FUSED_BS=128 UPSP_FUSED_OCC=12
run UPSP_FUSED_BS=64 UPSP_FUSED_OCC=16
run UPSP_FUSED_BS=64 UPSP_FUSED_OCC=20
run UPSP_FUSED_BS=64 UPSP_FUSED_OCC=24
run UPSP_FUSED_BS=256 UPSP_FUSED_OCC=4
run UPSP_FUSED_BS=256 UPSP_FUSED_OCC=5
EXTRA="--batch 32" run UPSP_FUSED_BS=128 UPSP_FUSED_OCC=10
EXTRA="--batch 64" run UPSP_FUSED_BS=128 UPSP_FUSED_OCC=10
EXTRA="--batch 256" run UPSP_FUSED_BS=128 UPSP_FUSED_OCC=10
