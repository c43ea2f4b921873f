mkdir -p gpurun_out
O=gpurun_out/r2r
N=${1:-8}
( timeout 600 python -m pytest tests/test_multigpu.py -q -x --timeout 500 > ${O}_pytest.log 2>&1; echo "pytest multigpu rc=$?" )
tail -2 ${O}_pytest.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 2 --warmup 2 --e2e-steps 0 --cpu-seconds 0 --check > ${O}_bench_$n.json 2> ${O}_bench_$n.err
  echo "bench $n rc=$?"
  python -c "
import json
d=json.loads(open('${O}_bench_$n.json').read().strip().splitlines()[-1])
print('$n: ms/step', d['ms_per_step'], 'value', d['value'], d['stage_ms'], 'parity', d.get('parity_checked'), {k:v['mean_ms'] for k,v in d['kernels'].items()})
"
}
run ship_sm1 UPSP_SHIP=sm UPSP_STAGED_PEERS=7 UPSP_SHIP_BPSM=1
run ship_sm2 UPSP_SHIP=sm UPSP_STAGED_PEERS=7 UPSP_SHIP_BPSM=2
run ship_sm4 UPSP_SHIP=sm UPSP_STAGED_PEERS=7 UPSP_SHIP_BPSM=4
run ship_ce UPSP_STAGED_PEERS=7
