#!/bin/bash
# dev tool (gpurun --gpus 8): BASELINE config 4 (4096 running-barrel-roll problems over 8 GPUs) and config 5 (horizon x batch sweep at 8 GPUs) at HEAD
R=${1:-r02v}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus 8 --workload barrel --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_strong_barrel_8gpu.json 2> gpurun_out/${R}_strong_barrel_8gpu.err
tail -n 1 gpurun_out/${R}_strong_barrel_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('barrel', d['n_gpus'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'host', round((d.get('e2e_host_collect') or {}).get('value',0),1), 'ms', round(d['ms_per_step'],2))"
python tools/sweep.py --gpus 8 --batches 256,1024,4096,16384 > gpurun_out/${R}_sweep_8gpu.jsonl 2> gpurun_out/${R}_sweep_8gpu.err
grep -c "^{" gpurun_out/${R}_sweep_8gpu.jsonl
