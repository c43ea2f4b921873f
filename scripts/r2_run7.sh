mkdir -p gpurun_out
O=gpurun_out/r2l
( timeout 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -k "tma or chain or ring" > ${O}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -2 ${O}_pytest.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python scripts/r2_timeline.py 4096 > ${O}_tl_$n.log 2>&1
  echo "== $n: $(grep -E 'mode=|overlap' ${O}_tl_$n.log | tr '\n' ' ')"
  grep -E "decode |patch |project " ${O}_tl_$n.log
}
run tma16_np UPSP_PROJ=tma16 UPSP_PIPELINE=0
run tma16_serial UPSP_PROJ=tma16 UPSP_FRONT=serial
run tma16_ovl UPSP_PROJ=tma16
run tma12_serial UPSP_PROJ=tma12 UPSP_FRONT=serial UPSP_SCAN_BPSM=8
run tma12_ovl UPSP_PROJ=tma12 UPSP_SCAN_BPSM=1 UPSP_SCAN_THREADS=128
run tma12_ovl2 UPSP_PROJ=tma12
