#!/bin/bash
# round-2 GPU session E: find the nondeterminism at bench scale
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -rfE --tb=line -k "linear_attention_block_fused" 2>&1 | tail -15
for cfg in "" "SRGD_LA_SERIAL=1" "SRGD_CONV_SPLITK=0" "SRGD_LA_SERIAL=1 SRGD_CONV_SPLITK=0"; do
  echo "=== env: $cfg"
  env $cfg python -m pytest tests/test_gpu_shapes.py -m gpu -q --tb=line -k "batch16_is_deterministic" 2>&1 | tail -3
done
