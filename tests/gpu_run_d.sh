#!/bin/bash
# round-2 GPU session D: LA merge/epilogue, gn_apply variants, split-K off in invariant mode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -rfEP --tb=short > gpurun_out/r02_pytest_d1.log 2>&1; tail -3 gpurun_out/r02_pytest_d1.log
python -m pytest tests/test_gpu_unet.py tests/test_gpu_shapes.py -m gpu -q -rfEP --tb=short -k "eps_vs or teacher_forced_vs_reference or deterministic or bench_shape or reload or debug_conv" > gpurun_out/r02_pytest_d2.log 2>&1; tail -3 gpurun_out/r02_pytest_d2.log
for b in 16 1 8; do
  python bench.py --batch $b --steps 60 --warmup 5 --no_cpu_baseline --no_gpu_eager --dump_launches gpurun_out/r02_launches_d_b$b.txt > gpurun_out/r02_bench_d_batch$b.json 2> gpurun_out/r02_bench_d_batch$b.err
  head -c 300 gpurun_out/r02_bench_d_batch$b.json; echo
done
