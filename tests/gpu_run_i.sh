#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_edm.py -m gpu -q -rfEP --tb=short > gpurun_out/r02_pytest_i.log 2>&1; tail -30 gpurun_out/r02_pytest_i.log | cut -c1-200
