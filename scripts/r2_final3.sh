# round-2 last single-GPU validation: the driver's own commands
mkdir -p gpurun_out
O=gpurun_out/r2x
( timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 > ${O}_pytest.log 2>&1; echo "pytest -m gpu rc=$?" )
tail -4 ${O}_pytest.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?" ); tail -2 ${O}_smoke.log
timeout 400 python bench.py > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('${O}_bench_1gpu.json').read().strip().splitlines()[-1])
print('bench_1gpu', d.get('value'), d.get('ms_per_step'), d.get('stage_ms'), 'chain', (d.get('chain') or {}).get('frac_of_peak'), 'parity', d.get('parity_checked'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
PY
