mkdir -p gpurun_out
O=gpurun_out/r2o
for r in dflt 3.2.7 3.1.7 3.2.6 4.3.6 5.3.5 6.4.5 4.2.5 2.1.8; do
  if [ $r = dflt ]; then e=""; else e="UPSP_TMA_RING=$r"; fi
  env $e timeout 200 python scripts/r2_timeline.py 8192 > ${O}_tl_$r.log 2>&1
  echo "== ring $r: $(grep -E 'mode=' ${O}_tl_$r.log) | $(grep -E 'project ' ${O}_tl_$r.log)"
done
for r in dflt 3.2.7; do
  if [ $r = dflt ]; then e=""; else e="UPSP_TMA_RING=$r"; fi
  env $e UPSP_FRONT=serial UPSP_SCAN_BPSM=8 UPSP_SCAN_THREADS=256 timeout 200 python scripts/r2_timeline.py 8192 > ${O}_tls_$r.log 2>&1
  echo "== serial ring $r: $(grep -E 'mode=' ${O}_tls_$r.log) | $(grep -E 'decode |project ' ${O}_tls_$r.log | tr '\n' ' ')"
done
# ncu captures for profiles/: projection (packed TMA mode), scan, phase 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_project_tma|k_hot_scan12|k_phase2_sym" -s 6 -c 3 -f -o ${O}_prof python bench.py --frames 2048 --steps 1 --warmup 1 --e2e-steps 0 --cpu-seconds 0 > ${O}_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"k_phase2_sym" -c 1 -f -o ${O}_prof_p2 python bench.py --frames 20000 --steps 1 --warmup 0 --e2e-steps 0 --cpu-seconds 0 > ${O}_ncu2.log 2>&1
echo "ncu p2 rc=$?"
ls -la gpurun_out/r2o*.ncu-rep
