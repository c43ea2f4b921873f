mkdir -p gpurun_out
O=gpurun_out/r2ad
N=${1:-8}
run() { n=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 > ${O}_bench_$n.json 2> ${O}_bench_$n.err
  echo "bench $n rc=$?"
  python -c "
import json
d=json.loads(open('${O}_bench_$n.json').read().strip().splitlines()[-1])
print('$n: ms/step', d['ms_per_step'], 'value', d['value'], d['stage_ms'], 'parity', d.get('parity_checked'), {k:v['mean_ms'] for k,v in d['kernels'].items()})
"
}
run seg64 UPSP_FORCE_SEG64=1
run dflt UPSP_NOP=1
