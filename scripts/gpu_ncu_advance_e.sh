#!/usr/bin/env bash
# Round 2, GPU call W (1 GPU): standard (material) advance_e -- timing at 512^3 and 1024^3, one ncu --set full capture at 512^3
set -u
mkdir -p gpurun_out
S=gpurun_out/r2w_summary.txt
: > $S
timeout 200 python scripts/fields_std.py 512 2>&1 | tail -1 | tee -a $S
timeout 300 python scripts/fields_std.py 1024 2>&1 | tail -1 | tee -a $S
timeout 500 ncu --set full --clock-control none --import-source on -k regex:advance_e_kernel -s 2 -c 1 -o gpurun_out/r2w_advance_e_std -f \
    python scripts/fields_std.py 512 3 > gpurun_out/r2w_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $S
python scripts/ncu_summary.py gpurun_out/r2w_advance_e_std.ncu-rep > gpurun_out/r2w_512_advance_e_standard.txt 2>&1
head -60 gpurun_out/r2w_512_advance_e_standard.txt | tee -a $S
