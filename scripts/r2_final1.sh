# round-2 single-GPU records for profiles/
mkdir -p gpurun_out
O=gpurun_out/r2v
timeout 900 python bench.py --steps 5 --warmup 3 > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_reference_arm.json 2> ${O}_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches_frames2048.csv python bench.py --frames 2048 --steps 2 --warmup 1 --e2e-steps 0 --cpu-seconds 0 --no-check > ${O}_launch_list.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"k_project_tma|k_hot_scan12|k_patch" -s 8 -c 4 -f -o /tmp/r2v_prof_phase1 python bench.py --frames 2048 --steps 1 --warmup 1 --e2e-steps 0 --cpu-seconds 0 --no-check > ${O}_ncu1.log 2>&1; echo "ncu phase1 rc=$?"
python scripts/ncu_summary.py /tmp/r2v_prof_phase1.ncu-rep ${O}_ncu_full_phase1.json
timeout 600 ncu --set full --clock-control none -k regex:"k_phase2_sym" -c 1 -f -o /tmp/r2v_prof_phase2 python bench.py --steps 1 --warmup 0 --e2e-steps 0 --cpu-seconds 0 --no-check > ${O}_ncu2.log 2>&1; echo "ncu phase2 rc=$?"
python scripts/ncu_summary.py /tmp/r2v_prof_phase2.ncu-rep ${O}_ncu_full_phase2.json
for c in 2 3; do
timeout 900 python bench.py --config $c --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 10 > ${O}_bench_config$c.json 2> ${O}_bench_config$c.err; echo "config $c rc=$?"
done
timeout 900 python bench.py --registration pixel --frames 4096 --steps 2 --warmup 2 --e2e-steps 0 --cpu-seconds 0 --no-check > ${O}_bench_regpixel.json 2> ${O}_bench_regpixel.err; echo "registration pixel rc=$?"
timeout 600 python bench.py --csr random --steps 3 --warmup 3 --e2e-steps 0 --cpu-seconds 0 > ${O}_bench_csr_random.json 2> ${O}_bench_csr_random.err; echo "csr random rc=$?"
timeout 900 python scripts/sweep_projection.py --out ${O}_projection_sweep.csv > ${O}_sweep.log 2>&1; echo "sweep rc=$?"
for f in bench_1gpu bench_reference_arm bench_config2 bench_config3 bench_regpixel bench_csr_random; do
python - <<PY
import json
try:
    d=json.loads(open('${O}_$f.json').read().strip().splitlines()[-1])
    print('$f', d.get('value'), d.get('ms_per_step'), d.get('stage_ms'), 'chain', (d.get('chain') or {}).get('frac_of_peak'), 'parity', d.get('parity_checked'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('cores'))
except Exception as e:
    print('$f', 'ERR', e)
PY
done
tail -3 ${O}_sweep.log
