#!/bin/bash
# dev tool, run on the GPU box: ncu launch list of one 4096-problem MHPC trot solve + one --set full capture per hot kernel.
# Each ncu pass only runs after the same command exited 0 without ncu. Outputs land in gpurun_out/ (summarised by tools/summarize_profiles.py).
set -e
R=${1:-r01}
CMD="python tools/profile_cmd.py mhpc 4096"
$CMD > gpurun_out/${R}_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_mhpc.csv $CMD > gpurun_out/${R}_ncu_launch.log 2>&1
# the full captures run with one stream per tick (CAFE_SPLIT_MIN=0): each captured launch then covers the whole active list, like the
# per-launch times bench.py measures in profiling mode (the launch list above keeps the default two-stream ticks)
export CAFE_SPLIT_MIN=0
CMD2="python tools/profile_cmd.py mhpc 4096 1 3"
$CMD2 > gpurun_out/${R}_plain2.log 2>&1
ncu --set full --clock-control none -k regex:'k_bwd2|k_lq|k_roll' -s 4 -c 5 -o /tmp/${R}_full_mhpc -f $CMD2 > gpurun_out/${R}_ncu_full.log 2>&1
# the report itself is ~100 MB (SASS of the generated whole-body code): only its raw metric page travels back
ncu -i /tmp/${R}_full_mhpc.ncu-rep --page raw --csv > gpurun_out/${R}_full_mhpc_raw.csv
tail -n 3 gpurun_out/${R}_plain.log
