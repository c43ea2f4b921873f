mkdir -p gpurun_out
O=gpurun_out/r2p
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > ${O}_smi.txt 2>&1
nvidia-smi topo -m >> ${O}_smi.txt 2>&1
( timeout 900 python -m pytest tests/test_multigpu.py tests/test_gpu_scale.py -q -x --timeout 600 -k "two_ranks or torchrun" > ${O}_pytest.log 2>&1; echo "pytest 2gpu rc=$?" )
tail -3 ${O}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 1 --cpu-seconds 0 --check > ${O}_bench2.json 2> ${O}_bench2.err
echo "bench 2gpu rc=$?"
grep -E "parity check|host affinity" ${O}_bench2.err | cut -c1-700
python -c "
import json
d=json.loads(open('${O}_bench2.json').read().strip().splitlines()[-1])
print('2gpu ms/step', d['ms_per_step'], 'value', d['value'], d['stage_ms'], 'parity', d.get('parity_checked'), 'e2e', d['e2e'] and d['e2e']['value'])
"
