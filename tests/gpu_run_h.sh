#!/bin/bash
# round-2 GPU session H: class-guidance sharing (tests + A/B), final single-GPU bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rfEP --tb=short > gpurun_out/r02_pytest_h.log 2>&1; tail -4 gpurun_out/r02_pytest_h.log
python bench.py --workload cfg32 --steps 30 --warmup 4 --no_cpu_baseline --no_gpu_eager > gpurun_out/r02_bench_h_cfg32.json 2> gpurun_out/r02_bench_h_cfg32.err
SRGD_CFG_SHARE=0 python bench.py --workload cfg32 --steps 30 --warmup 4 --no_cpu_baseline --no_gpu_eager > gpurun_out/r02_bench_h_cfg32_noshare.json 2> gpurun_out/r02_bench_h_cfg32_noshare.err
python bench.py --batch 1 --class_cond_scale 3.0 --steps 100 --warmup 5 --no_cpu_baseline --no_gpu_eager > gpurun_out/r02_bench_h_cfg1.json 2> gpurun_out/r02_bench_h_cfg1.err
SRGD_CFG_SHARE=0 python bench.py --batch 1 --class_cond_scale 3.0 --steps 100 --warmup 5 --no_cpu_baseline --no_gpu_eager > gpurun_out/r02_bench_h_cfg1_noshare.json 2> gpurun_out/r02_bench_h_cfg1_noshare.err
for f in gpurun_out/r02_bench_h_*.json; do echo $f; head -c 330 $f; echo; done
