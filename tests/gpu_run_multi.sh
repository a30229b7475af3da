#!/bin/bash
# round-2 multi-GPU session: run with `gpurun --gpus N -- bash tests/gpu_run_multi.sh N`
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if true; then
  $TR tests/gpu_multi_check.py > gpurun_out/r02_multi_check_n$N.log 2>&1; tail -5 gpurun_out/r02_multi_check_n$N.log
fi
$TR bench.py --gpus $N --workload tiled512 --steps 20 --warmup 4 --no_cpu_baseline > gpurun_out/r02_bench_tiled512_n$N.json 2> gpurun_out/r02_bench_tiled512_n$N.err
head -c 400 gpurun_out/r02_bench_tiled512_n$N.json; echo; tail -3 gpurun_out/r02_bench_tiled512_n$N.err
$TR bench.py --gpus $N --steps 60 --warmup 5 --no_cpu_baseline > gpurun_out/r02_bench_sample16_n$N.json 2> gpurun_out/r02_bench_sample16_n$N.err
head -c 400 gpurun_out/r02_bench_sample16_n$N.json; echo; tail -3 gpurun_out/r02_bench_sample16_n$N.err
head -c 400 gpurun_out/r02_bench_sample16_n$N.json; echo
$TR bench.py --gpus $N --workload tiled128 --steps 8 --warmup 2 --no_cpu_baseline > gpurun_out/r02_bench_tiled128_n$N.json 2> gpurun_out/r02_bench_tiled128_n$N.err
head -c 400 gpurun_out/r02_bench_tiled128_n$N.json; echo; tail -3 gpurun_out/r02_bench_tiled128_n$N.err
