mkdir -p gpurun_out
O=gpurun_out/r2q${N:-}
N=${1:-8}
nvidia-smi topo -m > ${O}_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 3 --warmup 3 --e2e-steps 1 --cpu-seconds 0 --check > ${O}_bench$N.json 2> ${O}_bench$N.err
echo "bench ${N}gpu rc=$?"
grep -E "parity check" ${O}_bench$N.err | cut -c1-330
python -c "
import json
d=json.loads(open('${O}_bench$N.json').read().strip().splitlines()[-1])
print('${N}gpu ms/step', d['ms_per_step'], 'value', d['value'], d['stage_ms'], 'parity', d.get('parity_checked'), 'e2e', d['e2e'] and d['e2e']['value'], {k:v['mean_ms'] for k,v in d['kernels'].items()})
"
