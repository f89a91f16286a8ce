#!/bin/bash
# dev tool, run on a multi-GPU box (gpurun --gpus 8): strong-scaling lines of the headline workload (BASELINE metric: 4096 MHPC trot problems
# over 1/2/4/8 GPUs), BASELINE config 4 (4096 running-barrel-roll problems over 2/4/8 GPUs) and config 5 (horizon x batch sweep at 8 GPUs).
R=${1:-r02}
NG=$(nvidia-smi -L | wc -l)
run() {  # n, extra args, tag
  local n=$1 tag=$3
  if [ "$n" = 1 ]; then python bench.py --gpus 1 $2 > gpurun_out/${R}_${tag}_1gpu.json 2> gpurun_out/${R}_${tag}_1gpu.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n $2 > gpurun_out/${R}_${tag}_${n}gpu.json 2> gpurun_out/${R}_${tag}_${n}gpu.err; fi
  tail -n 1 gpurun_out/${R}_${tag}_${n}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$tag', d['n_gpus'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2))"
}
for n in 1 2 4 8; do [ $n -le $NG ] && run $n "--steps 5 --warmup 3 --no-cpu-baseline" strong_mhpc; done
for n in 2 4 8; do [ $n -le $NG ] && run $n "--workload barrel --steps 2 --warmup 3 --no-cpu-baseline" strong_barrel; done
[ $NG -ge 2 ] && python tools/sweep.py --gpus $NG --batches 256,1024,4096,16384 > gpurun_out/${R}_sweep_${NG}gpu.jsonl 2> gpurun_out/${R}_sweep_${NG}gpu.err
tail -n 3 gpurun_out/${R}_sweep_${NG}gpu.jsonl
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
