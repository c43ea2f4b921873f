mkdir -p gpurun_out
O=gpurun_out/r2al
( timeout 1200 python -m pytest tests/test_multigpu.py tests/test_gpu_parity.py tests/test_gpu_scale.py -q -x --timeout 900 -k "blocked or ranks or 16bit or clustered or phase2 or chain or baseline" > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -12 ${O}_pytest.log
for v in plain blocked; do
  if [ $v = blocked ]; then export UPSP_BLOCKED=1; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --e2e-steps 1 --cpu-seconds 0 > ${O}_bench_$v.json 2> ${O}_bench_$v.err; echo "bench $v rc=$?"
  python -c "
import json
d=json.loads(open('${O}_bench_$v.json').read().strip().splitlines()[-1])
print('$v 1gpu ms/step', d['ms_per_step'], d['stage_ms'], d.get('parity_checked'), 'e2e', d['e2e'] and d['e2e']['value'])
"
done
