#!/bin/bash
# round-2 final validation (one GPU): the whole -m gpu suite, then compute-sanitizer memcheck over every kernel family
# (racecheck: see profiles/r02_compute_sanitizer.txt -- the continuous path and the EDM kernels report 0 hazards; the tool's
# host process is killed when the EDM block runs after them, so it is run per block: SAN_EDM / SAN_GAUSS / SAN_ONLY_EDM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rfEP --tb=short > gpurun_out/r02_pytest_final.log 2>&1; tail -4 gpurun_out/r02_pytest_final.log
timeout 900 compute-sanitizer --tool memcheck python tests/gpu_sanitize.py > gpurun_out/r02_sanitize_memcheck.log 2>&1
tail -5 gpurun_out/r02_sanitize_memcheck.log
SAN_EDM=0 SAN_ONLY_EDM=1 timeout 600 compute-sanitizer --tool racecheck python tests/gpu_sanitize.py > gpurun_out/r02_sanitize_racecheck_gauss.log 2>&1
tail -4 gpurun_out/r02_sanitize_racecheck_gauss.log
