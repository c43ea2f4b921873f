mkdir -p gpurun_out
O=gpurun_out/r2f
( timeout 300 python -m pytest tests -q -m gpu -x --timeout 240 > ${O}_pytest.log 2>&1; echo "pytest rc=$?" ) 
tail -3 ${O}_pytest.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python scripts/r2_timeline.py 4096 > ${O}_tl_$n.log 2>&1
  echo "== $n: $(grep -E 'mode=|overlap' ${O}_tl_$n.log | tr '\n' ' ')"
  grep -E "decode |patch |project " ${O}_tl_$n.log
}
run v4_base UPSP_PROJ=v4
run v4_p7d1 UPSP_PROJ=v4 UPSP_PROJ_PAD=12000 UPSP_DECODE_BPSM=1
run v4_p6d1 UPSP_PROJ=v4 UPSP_PROJ_PAD=15000 UPSP_DECODE_BPSM=1
run v4_p5d2 UPSP_PROJ=v4 UPSP_PROJ_PAD=20500 UPSP_DECODE_BPSM=2
run v4_p4d2 UPSP_PROJ=v4 UPSP_PROJ_PAD=28000 UPSP_DECODE_BPSM=2
run v4_nopipe UPSP_PROJ=v4 UPSP_PIPELINE=0
run tma12 UPSP_PROJ=tma12
run tma16 UPSP_PROJ=tma16
