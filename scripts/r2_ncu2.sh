mkdir -p gpurun_out
for m in tma12 tma16; do
UPSP_PROJ=$m UPSP_PIPELINE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_project_tma -s 3 -c 1 -f -o gpurun_out/r2i_prof_$m python bench.py --frames 2048 --steps 1 --warmup 1 --e2e-steps 0 --cpu-seconds 0 > gpurun_out/r2i_ncu_$m.log 2>&1
echo "ncu $m rc=$?"
done
ls -la gpurun_out/r2i*.ncu-rep
