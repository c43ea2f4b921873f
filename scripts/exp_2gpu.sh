i=0
for x in "--batch 256" "--batch 512" "--batch 128"; do
i=$((i+1))
echo "$x"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$i bench.py --gpus 2 --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0 $x 2>gpurun_out/b2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()})"
done
