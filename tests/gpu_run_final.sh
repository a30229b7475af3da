#!/bin/bash
# round-2 final single-GPU session: the driver's own commands first, then every workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_final_reference.json 2> gpurun_out/r02_bench_final_reference.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_final_driver_cmd.json 2> gpurun_out/r02_bench_final_driver_cmd.err
python bench.py --dump_launches gpurun_out/r02_launches_final_b16.txt > gpurun_out/r02_bench_final_sample16.json 2> gpurun_out/r02_bench_final_sample16.err
for wl in cfg32 tiled512 tiled128 sweep128; do
  python bench.py --workload $wl --steps 12 --warmup 3 > gpurun_out/r02_bench_final_$wl.json 2> gpurun_out/r02_bench_final_$wl.err
done
for b in 1 8; do
  python bench.py --batch $b --steps 100 --warmup 5 --no_cpu_baseline --no_gpu_eager --dump_launches gpurun_out/r02_launches_final_b$b.txt > gpurun_out/r02_bench_final_batch$b.json 2> gpurun_out/r02_bench_final_batch$b.err
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
for f in gpurun_out/r02_bench_final_*.json; do echo $f; tail -n 1 $f | head -c 330; echo; done
