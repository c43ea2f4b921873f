mkdir -p gpurun_out
O=gpurun_out/r2aj
run() { n=$1; shift
timeout 400 python bench.py --steps 2 --warmup 2 --e2e-steps 0 --cpu-seconds 0 --no-check "$@" > ${O}_bench_$n.json 2> ${O}_bench_$n.err; echo "bench $n rc=$?"
python -c "
import json
d=json.loads(open('${O}_bench_$n.json').read().strip().splitlines()[-1])
nb=-(-d['config']['frames_total']//256)
print('$n ms/step', d['ms_per_step'], d['stage_ms'], 'per batch', round(d['stage_ms']['process_frames']/nb,4), {k:v['mean_ms'] for k,v in d['kernels'].items()})
"
}
run n250k_f10000 --nodes 250000 --frames 10000
run n250k_f20000 --nodes 250000 --frames 20000
run n250k_f40000 --nodes 250000 --frames 40000
