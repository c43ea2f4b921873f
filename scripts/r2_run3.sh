mkdir -p gpurun_out
O=gpurun_out/r2g
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 240 -k "tma or hot" > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -2 ${O}_pytest.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python scripts/r2_timeline.py 4096 > ${O}_tl_$n.log 2>&1
  echo "== $n: $(grep -E 'mode=|overlap' ${O}_tl_$n.log | tr '\n' ' ')"
  grep -E "decode |patch |project " ${O}_tl_$n.log
}
run b2t256 UPSP_PROJ=tma12
run b1t256 UPSP_PROJ=tma12 UPSP_SCAN_BPSM=1
run b1t128 UPSP_PROJ=tma12 UPSP_SCAN_BPSM=1 UPSP_SCAN_THREADS=128
run b2t128 UPSP_PROJ=tma12 UPSP_SCAN_BPSM=2 UPSP_SCAN_THREADS=128
run b1t64 UPSP_PROJ=tma12 UPSP_SCAN_BPSM=1 UPSP_SCAN_THREADS=64
run b4t64 UPSP_PROJ=tma12 UPSP_SCAN_BPSM=4 UPSP_SCAN_THREADS=64
run nopipe UPSP_PROJ=tma12 UPSP_PIPELINE=0
