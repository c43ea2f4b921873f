mkdir -p gpurun_out
for v in 0 1 2; do
UPSP_PROJ=tma16 UPSP_TMA_VAR=$v timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 > gpurun_out/r2d_bench_var$v.json 2> gpurun_out/r2d_bench_var$v.err
echo "var $v rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2d_bench_var$v.json'))
print('var$v', d['ms_per_step'], d['stage_ms']['process_frames'], {k:v['mean_ms'] for k,v in d['kernels'].items()})
"
done
