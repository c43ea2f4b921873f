mkdir -p gpurun_out
O=gpurun_out/r2ah
run() { n=$1; shift
env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 --no-check > ${O}_bench_$n.json 2> ${O}_bench_$n.err; echo "bench $n rc=$?"
python -c "
import json
d=json.loads(open('${O}_bench_$n.json').read().strip().splitlines()[-1])
print('$n ms/step', d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()})
"
}
run vmm UPSP_FORCE_VMM=1
run vmm_seg UPSP_FORCE_VMM=1 UPSP_FORCE_SEG128=1
run plain UPSP_NOP=1
