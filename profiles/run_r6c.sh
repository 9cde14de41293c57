#!/bin/bash
# 8 GPUs, the 30 000-series recursive forecast after the short-window spectrum work (3 750 series per GPU)
t=r6c
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29718 bench.py --gpus 8 --steps 3 --warmup 3 --workload recursive --no-cpu-baseline > gpurun_out/${t}_bench_recursive_8gpu.json 2> gpurun_out/${t}_bench_recursive_8gpu.err; echo "recursive 8 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r6c_bench_recursive_8gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), d.get('series_per_sec'), d.get('periods_identical_on_all_ranks'))"
