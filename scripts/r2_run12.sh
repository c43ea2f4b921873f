mkdir -p gpurun_out
O=gpurun_out/r2u
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -q -x --timeout 600 > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -4 ${O}_pytest.log
for v in 1 0; do
UPSP_PHASE2_STREAM=$v timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 --check > ${O}_bench_p$v.json 2> ${O}_bench_p$v.err; echo "bench stream=$v rc=$?"
python -c "
import json
d=json.loads(open('${O}_bench_p$v.json').read().strip().splitlines()[-1])
print('stream=$v ms/step', d['ms_per_step'], d['stage_ms'], d.get('parity_checked'), 'chain', d['chain']['frac_of_peak'])
"
done
