mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
for m in v4 tma16 tma12; do
  timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 180 --maxfail=6 -k "test_tma_projection_modes and $m" > gpurun_out/r2a_test_$m.log 2>&1
  echo "pytest $m rc=$?" >> gpurun_out/r2a_summary.txt
done
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 180 -k "weighted_values or variants" > gpurun_out/r2a_test_misc.log 2>&1
echo "pytest misc rc=$?" >> gpurun_out/r2a_summary.txt
for m in v4 tma16 tma12; do
  UPSP_PROJ=$m timeout 400 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 > gpurun_out/r2a_bench_$m.json 2> gpurun_out/r2a_bench_$m.err
  echo "bench $m rc=$?" >> gpurun_out/r2a_summary.txt
done
cat gpurun_out/r2a_summary.txt
tail -5 gpurun_out/r2a_test_*.log
