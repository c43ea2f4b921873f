#!/bin/bash
# Round-end measurement run on one B200: bench line, ncu launch list, ncu --set full summaries.
# (ncu reports are summarised on the box and deleted: gpurun_out/ may carry at most 64 MiB back.)
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --frames 2048 --steps 2 --warmup 1 --cpu-seconds 0 --e2e-steps 0 > gpurun_out/final_launches.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_project_fused4|k_unpack12_scan_p|k_patch" -c 6 \
    -o gpurun_out/final_prof_phase1 python bench.py --frames 1024 --steps 1 --warmup 1 --cpu-seconds 0 --e2e-steps 0 > gpurun_out/final_prof1.log 2>&1
python scripts/ncu_summary.py gpurun_out/final_prof_phase1.ncu-rep gpurun_out/final_ncu_phase1.json; rm -f gpurun_out/final_prof_phase1.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:"k_phase2_sym" -c 1 \
    -o gpurun_out/final_prof_phase2 python bench.py --steps 1 --warmup 0 --cpu-seconds 0 --e2e-steps 0 > gpurun_out/final_prof2.log 2>&1
python scripts/ncu_summary.py gpurun_out/final_prof_phase2.ncu-rep gpurun_out/final_ncu_phase2.json; rm -f gpurun_out/final_prof_phase2.ncu-rep
du -sh gpurun_out
cut -c1-300 gpurun_out/final_bench.json
