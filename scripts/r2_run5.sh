mkdir -p gpurun_out
O=gpurun_out/r2j
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python scripts/r2_timeline.py 4096 > ${O}_tl_$n.log 2>&1
  echo "== $n: $(grep -E 'mode=|overlap' ${O}_tl_$n.log | tr '\n' ' ')"
  grep -E "decode |patch |project " ${O}_tl_$n.log
}
run tma16_d1 UPSP_PROJ=tma16 UPSP_DECODE_BPSM=1
run tma16_d2 UPSP_PROJ=tma16 UPSP_DECODE_BPSM=2
run tma16_d3 UPSP_PROJ=tma16 UPSP_DECODE_BPSM=3
run tma16_np UPSP_PROJ=tma16 UPSP_DECODE_P=0
( timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -q -x -s --timeout 500 -k "scale or ring_lap" > ${O}_scale.log 2>&1; echo "pytest scale rc=$?" )
grep -E "delta-Cp|passed|failed" ${O}_scale.log
