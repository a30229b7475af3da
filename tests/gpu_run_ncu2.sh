#!/bin/bash
# ncu launch list + --set full captures of one 16-tile bench step at the final HEAD (one GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
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
ls -la gpurun_out/*.gz gpurun_out/r02_ncu_launches_b16.csv
