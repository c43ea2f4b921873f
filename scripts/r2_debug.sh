mkdir -p gpurun_out
which compute-sanitizer > gpurun_out/r2b_san.log 2>&1
UPSP_PROJ=tma16 timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q --timeout 800 -x -k "test_tma_projection_modes and one-batch-tma16" >> gpurun_out/r2b_san.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_san.log
grep -n "=========" gpurun_out/r2b_san.log | head -60
