mkdir -p gpurun_out
O=gpurun_out/r2y
( timeout 1200 python -m pytest tests -q -x -m gpu --timeout 600 > ${O}_pytest.log 2>&1; echo "pytest gpu rc=$?" )
tail -6 ${O}_pytest.log
python __graft_entry__.py smoke > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${O}_smoke.log
