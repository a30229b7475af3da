#!/bin/bash
# round-2 GPU session C: split-K conv
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -rfEP --tb=short -k "conv" > gpurun_out/r02_pytest_c1.log 2>&1; tail -3 gpurun_out/r02_pytest_c1.log
python -m pytest tests/test_gpu_unet.py tests/test_gpu_shapes.py -m gpu -q -rfEP --tb=short -k "eps_vs or teacher_forced_vs_reference or deterministic or (bench_shape and unit) or reload" > gpurun_out/r02_pytest_c2.log 2>&1; tail -3 gpurun_out/r02_pytest_c2.log
for b in 16 1; do
  python bench.py --batch $b --steps 60 --warmup 5 --no_cpu_baseline --no_gpu_eager --dump_launches gpurun_out/r02_launches_c_b$b.txt > gpurun_out/r02_bench_c_batch$b.json 2> gpurun_out/r02_bench_c_batch$b.err
  head -c 300 gpurun_out/r02_bench_c_batch$b.json; echo
  SRGD_CONV_SPLITK=0 python bench.py --batch $b --steps 60 --warmup 5 --no_cpu_baseline --no_gpu_eager > gpurun_out/r02_bench_c_batch${b}_nosplit.json 2> gpurun_out/r02_bench_c_batch${b}_nosplit.err
  head -c 300 gpurun_out/r02_bench_c_batch${b}_nosplit.json; echo
done
