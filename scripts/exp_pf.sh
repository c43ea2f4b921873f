run() {
  echo -n "$*: "
  env "$@" timeout 200 python bench.py --frames 6400 --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0 $EXTRA 2>gpurun_out/exp_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']['process_frames'], {k:v['mean_ms'] for k,v in d['kernels'].items()})"
}
run UPSP_PIPELINE=0 UPSP_FUSED_DBG=0
run UPSP_PIPELINE=0 UPSP_FUSED_DBG=8
run UPSP_PIPELINE=1 UPSP_FUSED_DBG=0
run UPSP_PIPELINE=1 UPSP_FUSED_DBG=8
UPSP_FUSED_DBG=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
