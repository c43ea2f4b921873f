mkdir -p gpurun_out
O=gpurun_out/r2ae
( timeout 1200 python -m pytest tests/test_multigpu.py tests/test_gpu_parity.py -q -x --timeout 900 -k "ranks or clustered or tma or 16bit" > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -4 ${O}_pytest.log
UPSP_FORCE_SEG128=1 timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 > ${O}_bench_seg.json 2> ${O}_bench_seg.err; echo "bench seg rc=$?"
timeout 300 python bench.py --config 2 --steps 2 --warmup 2 --e2e-steps 0 --cpu-seconds 0 --no-check > ${O}_bench_c2.json 2> ${O}_bench_c2.err; echo "bench c2 rc=$?"
for n in bench_seg bench_c2; do python -c "
import json
d=json.loads(open('${O}_$n.json').read().strip().splitlines()[-1])
print('$n ms/step', d['ms_per_step'], d['stage_ms'], d.get('parity_checked'), {k:v['mean_ms'] for k,v in d['kernels'].items()})
"; done
