# round-2 final single-GPU validation: the driver's own commands, then the records for profiles/
mkdir -p gpurun_out
O=gpurun_out/r2w
( timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 > ${O}_pytest.log 2>&1; echo "pytest -m gpu rc=$?" )
tail -4 ${O}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?" ); tail -2 ${O}_smoke.log
timeout 900 python bench.py > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"k_phase2_sym" -c 1 -f -o /tmp/r2w_prof_phase2 python bench.py --steps 1 --warmup 0 --e2e-steps 0 --cpu-seconds 0 --no-check > ${O}_ncu2.log 2>&1; echo "ncu phase2 rc=$?"
python scripts/ncu_summary.py /tmp/r2w_prof_phase2.ncu-rep ${O}_ncu_full_phase2.json
python - <<PY
import json
d=json.loads(open('${O}_bench_1gpu.json').read().strip().splitlines()[-1])
print('bench_1gpu', d.get('value'), d.get('ms_per_step'), d.get('stage_ms'), 'chain', (d.get('chain') or {}).get('frac_of_peak'), 'parity', d.get('parity_checked'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('cores'), 'roofline', d.get('roofline'))
PY
