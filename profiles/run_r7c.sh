#!/bin/bash
# 4 GPUs, default bench (elec), final binary
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29724 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r7c_bench_elec_4gpu.json 2> gpurun_out/r7c_bench_elec_4gpu.err; echo "rc=$?"
wc -l gpurun_out/r7c_bench_elec_4gpu.json
python -c "
import json
d=json.loads(open('gpurun_out/r7c_bench_elec_4gpu.json').read())
print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), d['e2e']['value'], d.get('periods_identical_on_all_ranks'))"
