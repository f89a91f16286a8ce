#!/bin/bash
# dev tool, run on the GPU box: ncu --set full capture (with source page) of the cooperative whole-body kernels
set -e
K=${1:-'k_wb_fwd|k_wb_lq'}
B=${2:-4096}
export CAFE_SPLIT_MIN=0
CMD="python tools/profile_cmd.py mhpc $B 1 3"
$CMD > gpurun_out/prof_plain.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"$K" -s 2 -c 2 -o /tmp/prof_wb -f $CMD > gpurun_out/prof_ncu.log 2>&1
ncu -i /tmp/prof_wb.ncu-rep --page raw --csv > gpurun_out/prof_wb_raw.csv
ncu -i /tmp/prof_wb.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_wb_src.csv 2>/dev/null || true
python tools/ncu_raw.py /tmp/prof_wb.ncu-rep 2>/dev/null | head -90 || true
