#!/bin/bash
# round-2 final validation (one GPU): the whole -m gpu suite, then compute-sanitizer memcheck / racecheck
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rfEP --tb=short > gpurun_out/r02_pytest_final.log 2>&1; tail -4 gpurun_out/r02_pytest_final.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tests/gpu_sanitize.py > gpurun_out/r02_sanitize_$tool.log 2>&1
  tail -4 gpurun_out/r02_sanitize_$tool.log
done
