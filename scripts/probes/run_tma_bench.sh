# build here first (the binary travels to the GPU box, git ignores it): nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probes/tma_bench scripts/probes/tma_bench.cu -lcuda
P=scripts/probes/tma_bench
{
$P 96 6 1 6 8 2000 2 1
$P 96 6 1 12 8 2000 2 1
$P 96 6 1 16 8 1000 2 1
$P 96 6 1 24 4 1000 2 1
$P 96 6 1 6 4 2000 2 4
$P 96 6 1 6 4 2000 2 4 1
$P 96 6 1 6 2 1000 2 8
$P 96 6 1 1 4 2000 2 4
$P 96 6 1 1 4 2000 2 4 1
$P 96 6 4 6 2 500 2 1
$P 96 6 4 6 2 500 2 4
$P 96 8 4 6 2 500 2 4
$P 96 6 1 6 1 2000 2 1
$P 96 6 1 6 2 2000 2 1
$P 96 6 1 6 4 2000 2 1
} > gpurun_out/r2e_tma_bench2.log 2>&1
cat gpurun_out/r2e_tma_bench2.log
