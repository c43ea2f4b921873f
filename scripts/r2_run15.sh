mkdir -p gpurun_out
O=gpurun_out/r2aa
run() { n=$1; shift
env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 --no-check > ${O}_bench_$n.json 2> ${O}_bench_$n.err; echo "bench $n rc=$?"
python -c "
import json
d=json.loads(open('${O}_bench_$n.json').read().strip().splitlines()[-1])
print('$n ms/step', d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()})
"
}
run seg_serial8 UPSP_FORCE_SEG128=1 UPSP_FRONT=serial UPSP_SCAN_BPSM=8 UPSP_SCAN_THREADS=256
run seg_serial4 UPSP_FORCE_SEG128=1 UPSP_FRONT=serial UPSP_SCAN_BPSM=4 UPSP_SCAN_THREADS=256
run seg_ovl_b2t64 UPSP_FORCE_SEG128=1 UPSP_SCAN_BPSM=1 UPSP_SCAN_THREADS=64
run seg_ring325 UPSP_FORCE_SEG128=1 UPSP_TMA_SPLIT=4
run n1_serial8 UPSP_FRONT=serial UPSP_SCAN_BPSM=8 UPSP_SCAN_THREADS=256
