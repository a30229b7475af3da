#!/bin/bash
# end-of-round confirmation at the final HEAD (one GPU): whole -m gpu suite, smoke, the driver's two bench commands
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rfE --tb=short > gpurun_out/r02_pytest_final2.log 2>&1; tail -3 gpurun_out/r02_pytest_final2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke2.log 2>&1; tail -2 gpurun_out/r02_smoke2.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_final2_reference.json 2> gpurun_out/r02_bench_final2_reference.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_final2_sample16.json 2> gpurun_out/r02_bench_final2_sample16.err
python bench.py --batch 1 --steps 100 --warmup 5 --no_cpu_baseline --no_gpu_eager > gpurun_out/r02_bench_final2_batch1.json 2> gpurun_out/r02_bench_final2_batch1.err
for f in gpurun_out/r02_bench_final2_*.json; do echo $f; tail -n 1 $f | head -c 400; echo; done
