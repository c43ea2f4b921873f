mkdir -p gpurun_out
O=gpurun_out/r2ab
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -q -x --timeout 900 > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -8 ${O}_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"
UPSP_FORCE_SEG128=1 timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 > ${O}_bench_seg.json 2> ${O}_bench_seg.err; echo "bench seg rc=$?"
for n in bench bench_seg; do python -c "
import json
d=json.loads(open('${O}_$n.json').read().strip().splitlines()[-1])
print('$n ms/step', d['ms_per_step'], d['stage_ms'], d.get('parity_checked'), {k:v['mean_ms'] for k,v in d['kernels'].items()})
"; done
