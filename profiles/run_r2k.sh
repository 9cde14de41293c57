#!/bin/bash
# round 2, re-entry GPU pass: tests, smoke, bench lines for every workload
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2k_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/r2k_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_elec.json 2> gpurun_out/r2k_bench_elec.err; echo "bench rc=$?"; tail -3 gpurun_out/r2k_bench_elec.err
timeout 300 python bench.py --steps 3 --warmup 3 --workload recursive --no-cpu-baseline > gpurun_out/r2k_bench_recursive.json 2> gpurun_out/r2k_bench_recursive.err; echo "recursive rc=$?"; tail -3 gpurun_out/r2k_bench_recursive.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload etth1 --no-cpu-baseline > gpurun_out/r2k_bench_etth1.json 2> gpurun_out/r2k_bench_etth1.err; echo "etth1 rc=$?"; tail -3 gpurun_out/r2k_bench_etth1.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --no-cpu-baseline > gpurun_out/r2k_bench_traffic.json 2> gpurun_out/r2k_bench_traffic.err; echo "traffic rc=$?"; tail -3 gpurun_out/r2k_bench_traffic.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --dtype bf16 --no-cpu-baseline > gpurun_out/r2k_bench_traffic_bf16.json 2> gpurun_out/r2k_bench_traffic_bf16.err; echo "traffic bf16 rc=$?"; tail -3 gpurun_out/r2k_bench_traffic_bf16.err
head -c 1500 gpurun_out/r2k_bench_elec.json
