mkdir -p gpurun_out
O=gpurun_out/r2m
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py tests/test_gpu_scale.py -q -x --timeout 400 > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -2 ${O}_pytest.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python scripts/r2_timeline.py 4096 > ${O}_tl_$n.log 2>&1
  echo "== $n: $(grep -E 'mode=|overlap' ${O}_tl_$n.log | tr '\n' ' ')"
  grep -E "decode |patch |project " ${O}_tl_$n.log
}
run tma12_s1 UPSP_PROJ=tma12 UPSP_TMA_SPLIT=1
run tma12_s2 UPSP_PROJ=tma12 UPSP_TMA_SPLIT=2
run tma12_s4 UPSP_PROJ=tma12 UPSP_TMA_SPLIT=4
run tma16_np_s2 UPSP_PROJ=tma16 UPSP_PIPELINE=0 UPSP_TMA_SPLIT=2
run tma16_np_s4 UPSP_PROJ=tma16 UPSP_PIPELINE=0 UPSP_TMA_SPLIT=4
for b in 0 512; do
timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 --batch $b > ${O}_bench_b$b.json 2> ${O}_bench_b$b.err; echo "bench batch $b rc=$?"
python -c "
import json
d=json.loads(open('${O}_bench_b$b.json').read().strip().splitlines()[-1])
print('batch $b ms/step', d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()})
"
done
