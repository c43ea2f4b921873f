mkdir -p gpurun_out
O=gpurun_out/r2ao
run() {  # name cl nt extra-args...
  name=$1; cl=$2; nt=$3; shift 3
  UPSP_P2_CL=$cl UPSP_P2_NT=$nt timeout 300 python bench.py --steps 3 --warmup 2 --e2e-steps 0 --cpu-seconds 0 --no-check "$@" > ${O}_$name.json 2> ${O}_$name.err; rc=$?
  python -c "
import json
try:
    d=json.loads(open('${O}_$name.json').read().strip().splitlines()[-1])
    print('$name cl=$cl nt=$nt rc=$rc ms/step', d['ms_per_step'], 'phase2', d['stage_ms']['phase2'])
except Exception as e:
    print('$name cl=$cl nt=$nt rc=$rc ERR', e)
"
}
S="--width 512 --height 512"
run f80k_c4n512 4 512 --frames 80000 --nodes 125000 $S
run f80k_c8n256 8 256 --frames 80000 --nodes 125000 $S
run f160k_c8n512 8 512 --frames 160000 --nodes 62500 $S
run f160k_c8n256 8 256 --frames 160000 --nodes 62500 $S
export UPSP_BLOCKED=1
run f40k_blk_c2n512 2 512 --frames 40000 --nodes 250000 $S
run f40k_blk_c4n256 4 256 --frames 40000 --nodes 250000 $S
