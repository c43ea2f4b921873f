mkdir -p gpurun_out
O=gpurun_out/r2h
( timeout 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -k "tma or hot or chain" > ${O}_pytest.log 2>&1; echo "pytest tma rc=$?" )
tail -3 ${O}_pytest.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python scripts/r2_timeline.py 4096 > ${O}_tl_$n.log 2>&1
  echo "== $n: $(grep -E 'mode=|overlap' ${O}_tl_$n.log | tr '\n' ' ')"
  grep -E "decode |patch |project " ${O}_tl_$n.log
}
run tma12_nopipe_dflt UPSP_PROJ=tma12 UPSP_SCAN_BPSM=1 UPSP_SCAN_THREADS=128 UPSP_PIPELINE=0
run tma12_b1t128 UPSP_PROJ=tma12 UPSP_SCAN_BPSM=1 UPSP_SCAN_THREADS=128
run tma12_b2t256 UPSP_PROJ=tma12
run tma16_nopipe UPSP_PROJ=tma16 UPSP_PIPELINE=0
run tma16 UPSP_PROJ=tma16
( timeout 600 python -m pytest tests/test_gpu_scale.py -q -x -s --timeout 500 > ${O}_scale.log 2>&1; echo "pytest scale rc=$?" )
tail -5 ${O}_scale.log
timeout 300 python bench.py --steps 1 --warmup 1 --frames 2048 --e2e-steps 0 --cpu-seconds 0 --check > ${O}_check.json 2> ${O}_check.err; echo "bench check rc=$?"
tail -3 ${O}_check.err
