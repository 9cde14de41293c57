#!/bin/bash
# 8-GPU pass: bench over the peer mailbox (default transport), one line per N
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r3c_bench_elec_${n}gpu.json 2> gpurun_out/r3c_bench_elec_${n}gpu.err; echo "bench $n rc=$?"; tail -2 gpurun_out/r3c_bench_elec_${n}gpu.err
done
