#!/bin/bash
# dev tool (gpurun --gpus 8): the headline strong-scaling lines at N = 8 and 4 with both ways of collecting the records (e2e: NCCL gather to rank 0;
# e2e_host_collect: every rank's shard over its own PCIe link into one shared page-locked host buffer)
R=${1:-r02t}
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_strong_mhpc_${n}gpu.json 2> gpurun_out/${R}_strong_mhpc_${n}gpu.err
  tail -n 1 gpurun_out/${R}_strong_mhpc_${n}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e_host_collect', round(d['e2e_host_collect']['value'],1))"
done
