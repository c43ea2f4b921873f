mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${1:-2} --master-addr 127.0.0.1 --master-port 29577 scripts/r2_timeline_dist.py 4096 > gpurun_out/r2ai_tl.log 2>&1
echo "rc=$?"
grep -E "rank |scan|patch|project" gpurun_out/r2ai_tl.log | cut -c1-300
