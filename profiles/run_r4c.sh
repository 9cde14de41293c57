#!/bin/bash
# 8-GPU pass: every BASELINE configuration that names 8 GPUs -- elec (weak scaling over the peer mailbox, strong-scaling
# block), traffic fp32 (spectrum exchange across ranks) and the 30 000-series recursive forecast (3 750 series per GPU)
nvidia-smi -L | wc -l
t=r4c
run() { # name nproc args...
  local name=$1 n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n "$@" > gpurun_out/${t}_bench_${name}_${n}gpu.json 2> gpurun_out/${t}_bench_${name}_${n}gpu.err; echo "$name $n rc=$?"; tail -2 gpurun_out/${t}_bench_${name}_${n}gpu.err
}
run elec 8 --steps 20 --warmup 5
run traffic 8 --steps 5 --warmup 3 --workload traffic --no-cpu-baseline
run recursive 8 --steps 3 --warmup 3 --workload recursive --no-cpu-baseline
run elec 2 --steps 20 --warmup 5 --no-cpu-baseline
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r4c_bench_*gpu.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], d["n_gpus"], round(d["value"]), round(d["ms_per_step"], 4), d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
