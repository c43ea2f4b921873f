run() {
  echo -n "$*: "
  env "$@" timeout 200 python bench.py --frames 6400 --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0 $EXTRA 2>gpurun_out/exp_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']['process_frames'], {k:v['mean_ms'] for k,v in d['kernels'].items()})"
}
for o in 8 9 11 12; do run UPSP_PIPELINE=0 UPSP_FUSED_BS=128 UPSP_FUSED_OCC=$o; done
for o in 11 12; do run UPSP_PIPELINE=1 UPSP_FUSED_BS=128 UPSP_FUSED_OCC=$o; done
