mkdir -p gpurun_out
O=gpurun_out/r2ak
( timeout 1200 python -m pytest tests/test_multigpu.py tests/test_gpu_parity.py -q -x --timeout 900 -k "blocked or ranks or 16bit or clustered" > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -12 ${O}_pytest.log
UPSP_BLOCKED=1 timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-seconds 0 > ${O}_bench_blk.json 2> ${O}_bench_blk.err; echo "bench blocked rc=$?"
python -c "
import json
d=json.loads(open('${O}_bench_blk.json').read().strip().splitlines()[-1])
print('blocked 1gpu ms/step', d['ms_per_step'], d['stage_ms'], d.get('parity_checked'), 'e2e', d['e2e'] and d['e2e']['value'])
"
