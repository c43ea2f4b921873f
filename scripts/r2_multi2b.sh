mkdir -p gpurun_out
O=gpurun_out/r2x
( timeout 900 python -m pytest tests/test_gpu_scale.py -q -x --timeout 600 -k "torchrun" > ${O}_pytest.log 2>&1; echo "pytest torchrun rc=$?" )
tail -15 ${O}_pytest.log
for x in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 --exchange $x > ${O}_bench2_$x.json 2> ${O}_bench2_$x.err
echo "bench 2gpu $x rc=$?"
tail -2 ${O}_bench2_$x.err | cut -c1-300
python -c "
import json
d=json.loads(open('${O}_bench2_$x.json').read().strip().splitlines()[-1])
print('2gpu $x ms/step', d['ms_per_step'], 'value', d['value'], d['stage_ms'], 'parity', d.get('parity_checked'))
"
done
