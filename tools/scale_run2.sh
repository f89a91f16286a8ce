#!/bin/bash
# dev tool, run on an 8-GPU box (gpurun --gpus 8): strong-scaling lines of the headline workload at HEAD (4096 MHPC trot problems over 1/2/4/8 GPUs)
# and the multi-GPU parity tests.
R=${1:-r02s}
run() {
  local n=$1
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_strong_mhpc_1gpu.json 2> gpurun_out/${R}_strong_mhpc_1gpu.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_strong_mhpc_${n}gpu.json 2> gpurun_out/${R}_strong_mhpc_${n}gpu.err; fi
  tail -n 1 gpurun_out/${R}_strong_mhpc_${n}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('strong_mhpc', d['n_gpus'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2))"
}
for n in 8 4 2 1; do run $n; done
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>/dev/null; cut -c1-160 gpurun_out/${R}_bench_ref.json
