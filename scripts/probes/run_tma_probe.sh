# build here first (the binary travels to the GPU box, git ignores it): nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probes/tma_probe scripts/probes/tma_probe.cu -lcuda
P=scripts/probes/tma_probe
{
$P 128 96 150 80 6 0 0 0
$P 128 96 150 80 6 5 7 1
$P 128 96 150 80 6 -1 -1 0
$P 128 96 150 80 6 3 7 0
$P 128 96 150 80 6 8 8 0
$P 128 96 150 80 6 60 7 3
$P 128 96 150 80 6 121 93 2
$P 128 96 150 80 6 -200 5 2
$P 128 96 150 80 6 0 0 0 1024
$P 128 96 150 80 6 0 0 0 128
$P 128 96 150 80 6 0 0 0 64
$P 128 96 150 80 6 0 0 0 16
$P 1024 1024 16 80 6 333 777 3
$P 128 96 150 64 6 3 7 0
$P 128 96 150 128 6 3 7 0
$P 128 96 150 72 6 3 7 0
} > gpurun_out/r2c_tma_probe.log 2>&1
cat gpurun_out/r2c_tma_probe.log
