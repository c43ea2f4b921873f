mkdir -p gpurun_out
O=gpurun_out/r2k
( timeout 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -k "tma" > ${O}_pytest.log 2>&1; echo "pytest tma rc=$?" )
tail -2 ${O}_pytest.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python scripts/r2_timeline.py 4096 > ${O}_tl_$n.log 2>&1
  echo "== $n: $(grep -E 'mode=|overlap' ${O}_tl_$n.log | tr '\n' ' ')"
  grep -E "decode |patch |project " ${O}_tl_$n.log
}
run tma16_np_l0 UPSP_PROJ=tma16 UPSP_PIPELINE=0 UPSP_TMA_LANES=0
run tma16_np_l1 UPSP_PROJ=tma16 UPSP_PIPELINE=0 UPSP_TMA_LANES=1
run tma12_l0 UPSP_PROJ=tma12 UPSP_SCAN_BPSM=1 UPSP_SCAN_THREADS=128 UPSP_TMA_LANES=0
run tma12_l1 UPSP_PROJ=tma12 UPSP_SCAN_BPSM=1 UPSP_SCAN_THREADS=128 UPSP_TMA_LANES=1
