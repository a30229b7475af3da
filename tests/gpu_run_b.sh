#!/bin/bash
# round-2 GPU session B: shape tests + every bench workload at N=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_shapes.py -m gpu -q -rfEP --tb=short -k "bench_shape" > gpurun_out/r02_pytest_b.log 2>&1
tail -3 gpurun_out/r02_pytest_b.log
python bench.py --steps 20 --warmup 5 --dump_launches gpurun_out/r02_launches_b16.txt > gpurun_out/r02_bench_b_sample16.json 2> gpurun_out/r02_bench_b_sample16.err
for wl in cfg32 tiled512 tiled128 sweep128; do
  python bench.py --workload $wl --steps 8 --warmup 3 --no_cpu_baseline > gpurun_out/r02_bench_b_$wl.json 2> gpurun_out/r02_bench_b_$wl.err
  tail -c 300 gpurun_out/r02_bench_b_$wl.err
done
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_b_reference.json 2> gpurun_out/r02_bench_b_reference.err
for f in gpurun_out/r02_bench_b_*.json; do echo $f; head -c 400 $f; echo; done
for b in 1 8; do
  python bench.py --batch $b --steps 40 --warmup 5 --no_cpu_baseline --no_gpu_eager --dump_launches gpurun_out/r02_launches_b$b.txt > gpurun_out/r02_bench_b_batch$b.json 2> gpurun_out/r02_bench_b_batch$b.err
  head -c 300 gpurun_out/r02_bench_b_batch$b.json; echo
done
