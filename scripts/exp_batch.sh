for x in "--batch 128" "--batch 256" "--batch 512" "--batch 256"; do
echo "$x"
timeout 400 python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0 $x 2>gpurun_out/b1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'], {k:v['mean_ms'] for k,v in d['kernels'].items()})"
done
