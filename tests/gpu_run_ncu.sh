#!/bin/bash
# round-2 profiling session (one GPU): ncu launch list + --set full captures of one bench step at HEAD, CLI timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
git_head=$(cat .git_head 2>/dev/null || echo unknown)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_b16.csv python tests/gpu_profile_step.py --batch 16 --steps 1 > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm -f -o /tmp/r02_conv_full \
    python tests/gpu_profile_step.py --batch 16 --steps 1 > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
ncu -i /tmp/r02_conv_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_conv_raw.csv && gzip -f gpurun_out/r02_ncu_full_conv_raw.csv
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:gn_apply|la_|fa_tc|sampler|rmsnorm|pixel_inv|pack_input" -f -o /tmp/r02_misc_full \
    python tests/gpu_profile_step.py --batch 16 --steps 1 > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_c.log
ncu -i /tmp/r02_misc_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_misc_raw.csv && gzip -f gpurun_out/r02_ncu_full_misc_raw.csv
ls -la /tmp/*.ncu-rep gpurun_out | tail -12
python tests/gpu_cli_timing.py --images 40 > gpurun_out/r02_cli_timing.json 2> gpurun_out/r02_cli_timing.err
tail -30 gpurun_out/r02_cli_timing.json
