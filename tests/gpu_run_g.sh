#!/bin/bash
# round-2 GPU session G: everything after the LA / GN / split-K changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rfEP --tb=short > gpurun_out/r02_pytest_g.log 2>&1; tail -4 gpurun_out/r02_pytest_g.log
for b in 16 1 8; do
  python bench.py --batch $b --steps 60 --warmup 5 --no_cpu_baseline --no_gpu_eager --dump_launches gpurun_out/r02_launches_g_b$b.txt > gpurun_out/r02_bench_g_batch$b.json 2> gpurun_out/r02_bench_g_batch$b.err
  head -c 300 gpurun_out/r02_bench_g_batch$b.json; echo
done
