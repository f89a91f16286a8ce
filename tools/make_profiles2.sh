#!/bin/bash
# dev tool, run on the GPU box: round-2 profile set of the 4096-problem MHPC trot solve.
#  (1) ncu launch list (gpu__time_duration) of one whole solve with the default tick (two streams, overlapped linearisation),
#  (2) one `--set full --import-source on` capture of each hot kernel on a full-batch launch (first tick, single stream so that the
#      captured launch covers the whole 4096-problem list): k_wb_derivs, k_wb_sens, k_wb_cost, k_bwd2, k_wb_fwd.
# Each ncu pass runs only after the same command exited 0 without ncu. Raw and source pages travel back as CSV.
set -e
R=${1:-r02}
B=${2:-4096}
CMD="python tools/profile_cmd.py mhpc $B"
$CMD > gpurun_out/${R}_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches_mhpc.csv $CMD > gpurun_out/${R}_ncu_launch.log 2>&1
export CAFE_SPLIT_MIN=0 CAFE_LQ_OVERLAP=0
CMD2="python tools/profile_cmd.py mhpc $B 1 3"
$CMD2 > gpurun_out/${R}_plain2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'k_bwd2|k_wb_sens|k_wb_cost|k_wb_derivs|k_wb_fwd' -s 1 -c 5 -o /tmp/${R}_full -f $CMD2 > gpurun_out/${R}_ncu_full.log 2>&1
ncu -i /tmp/${R}_full.ncu-rep --page raw --csv > gpurun_out/${R}_full_raw.csv
ncu -i /tmp/${R}_full.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${R}_full_src.csv 2>/dev/null || true
ls -la /tmp/${R}_full.ncu-rep gpurun_out/${R}_full_src.csv
tail -n 3 gpurun_out/${R}_plain.log
