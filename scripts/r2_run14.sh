mkdir -p gpurun_out
O=gpurun_out/r2z
for v in 0 1; do
UPSP_FORCE_SEG128=$v timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 --no-check > ${O}_bench_seg$v.json 2> ${O}_bench_seg$v.err; echo "bench seg128=$v rc=$?"
python -c "
import json
d=json.loads(open('${O}_bench_seg$v.json').read().strip().splitlines()[-1])
print('seg128=$v ms/step', d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()})
"
done
