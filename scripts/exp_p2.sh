run() {
  echo -n "$*: "
  env "$@" timeout 250 python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0 $EXTRA 2>gpurun_out/exp_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()})"
}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run UPSP_PIPELINE=1
run UPSP_PIPELINE=0
