#!/bin/bash
cd "$(dirname "$0")/.."
for v in none res stage; do
  echo "=== variant $v"
  SRGD_B200_LIB=$PWD/srgd_b200/libsrgd_b200_$v.so python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=line -k "linear_attention_block_fused and 16-128-128-128" 2>&1 | tail -3
done
echo "=== default"
python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=line -k "linear_attention_block_fused and 16-128-128-128" 2>&1 | tail -3
