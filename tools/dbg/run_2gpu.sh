timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/chk_2gpu.json 2> gpurun_out/chk_2gpu.err; echo "rc=$?"
head -c 120 gpurun_out/chk_2gpu.json; echo; wc -l gpurun_out/chk_2gpu.json; grep -c "NCCL version" gpurun_out/chk_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/chk_2gpu_ref.json 2> gpurun_out/chk_2gpu_ref.err; echo "ref rc=$?"; head -c 200 gpurun_out/chk_2gpu_ref.json; echo
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=300 2>&1 | tail -3
