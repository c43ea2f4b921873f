mkdir -p gpurun_out
O=gpurun_out/r2s
( timeout 900 python -m pytest tests -q -x -m gpu --timeout 600 > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -4 ${O}_pytest.log
for v in 1 0; do
UPSP_ITRANS16=$v timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-seconds 0 --check > ${O}_bench_i$v.json 2> ${O}_bench_i$v.err; echo "bench it16=$v rc=$?"
python -c "
import json
d=json.loads(open('${O}_bench_i$v.json').read().strip().splitlines()[-1])
print('it16=$v ms/step', d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()}, d.get('parity_checked'), 'e2e', d['e2e'] and d['e2e']['value'], 'chain', d['chain']['frac_of_peak'])
"
done
