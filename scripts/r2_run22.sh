mkdir -p gpurun_out
O=gpurun_out/r2am
( UPSP_P2_VARIANT=7 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -q -x --timeout 600 -k "16bit or baseline" > ${O}_pytest.log 2>&1; echo "pytest (variant 7) rc=$?" )
tail -5 ${O}_pytest.log
for v in 0 1 2 3 6 7; do
  UPSP_P2_VARIANT=$v timeout 300 python bench.py --steps 4 --warmup 3 --e2e-steps 0 --cpu-seconds 0 > ${O}_bench_v$v.json 2> ${O}_bench_v$v.err; echo "bench v$v rc=$?"
  python -c "
import json
d=json.loads(open('${O}_bench_v$v.json').read().strip().splitlines()[-1])
print('v$v 1gpu ms/step', d['ms_per_step'], d['stage_ms'], d.get('parity_checked'))
"
done
