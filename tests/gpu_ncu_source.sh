#!/bin/bash
# ncu source-level captures of single launches of the bench step (developer tool; run under gpurun):
#   tests/gpu_ncu_source.sh "<name>:<kernel regex>:<launch-skip>" ...
# writes gpurun_out/src_<name>.csv.gz (ncu --page source, SASS) and gpurun_out/raw_<name>.csv.gz
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r name regex skip <<< "$spec"
  ncu --profile-from-start off --set full --import-source on --clock-control none -k "regex:$regex" \
      --launch-skip "$skip" --launch-count 1 -f -o /tmp/src_$name python tests/gpu_profile_step.py --batch 16 --steps 1 \
      > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/src_$name.ncu-rep --page source --csv --print-source sass > gpurun_out/src_$name.csv
  ncu -i /tmp/src_$name.ncu-rep --page raw --csv > gpurun_out/raw_$name.csv
  gzip -f gpurun_out/src_$name.csv gpurun_out/raw_$name.csv
done
ls -la gpurun_out | tail -20
