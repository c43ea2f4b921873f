i=0
for x in "UPSP_STAGED_PEERS=3" "UPSP_STAGED_PEERS=2"; do
i=$((i+1))
echo "$x"
env $x timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2956$i bench.py --gpus 8 --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0 2>gpurun_out/b8.err | tee gpurun_out/bench_8gpu_k$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()})"
done
