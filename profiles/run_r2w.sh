#!/bin/bash
# round 2: 2-GPU pass -- peer mailbox test, bench over both transports, 1-GPU line of the same box
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=300 -x > gpurun_out/r2w_multi.log 2>&1; echo "multi rc=$?"; tail -15 gpurun_out/r2w_multi.log
for t in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --transport $t > gpurun_out/r2w_bench_elec_2gpu_$t.json 2> gpurun_out/r2w_bench_elec_2gpu_$t.err; echo "bench $t rc=$?"; tail -2 gpurun_out/r2w_bench_elec_2gpu_$t.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2w_bench_elec_1gpu.json 2> gpurun_out/r2w_bench_elec_1gpu.err; echo "bench1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload recursive --no-cpu-baseline > gpurun_out/r2w_bench_recursive_2gpu.json 2> gpurun_out/r2w_bench_recursive_2gpu.err; echo "recursive2 rc=$?"; tail -2 gpurun_out/r2w_bench_recursive_2gpu.err
