run() {
  echo -n "$*: "
  env "$@" timeout 250 python bench.py --frames 6400 --steps 3 --warmup 3 --cpu-seconds 0 --e2e-steps 0 $EXTRA 2>gpurun_out/exp_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']['process_frames'], {k:v['mean_ms'] for k,v in d['kernels'].items()})"
}
for rep in 1 2; do
run UPSP_PIPELINE=0 UPSP_FUSED_DBG=0
run UPSP_PIPELINE=0 UPSP_FUSED_DBG=16
run UPSP_PIPELINE=0 UPSP_FUSED_V=3
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
