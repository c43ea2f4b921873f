mkdir -p gpurun_out
O=gpurun_out/r2n
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py tests/test_gpu_scale.py tests/test_host_driver.py -q -x --timeout 400 > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -3 ${O}_pytest.log
for v in 0 1; do
UPSP_PHASE2_SCALAR=$v timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 --check > ${O}_bench_s$v.json 2> ${O}_bench_s$v.err; echo "bench scalar=$v rc=$?"
python -c "
import json
d=json.loads(open('${O}_bench_s$v.json').read().strip().splitlines()[-1])
print('scalar=$v ms/step', d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()}, d.get('parity_checked'), d['parity']['ranks'][0])
"
done
