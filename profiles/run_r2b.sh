#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout=300 > gpurun_out/r2e_tc.log 2>&1; echo "tc rc=$?"; tail -15 gpurun_out/r2e_tc.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --deselect tests/test_gpu_tc.py > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2e_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --workload etth1 --no-cpu-baseline > gpurun_out/r2e_bench_etth1.json 2> gpurun_out/r2e_bench_etth1.err; echo "etth1 rc=$?"; tail -2 gpurun_out/r2e_bench_etth1.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --no-cpu-baseline > gpurun_out/r2e_bench_traffic.json 2> gpurun_out/r2e_bench_traffic.err; echo "traffic rc=$?"; tail -2 gpurun_out/r2e_bench_traffic.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --dtype bf16 --no-cpu-baseline > gpurun_out/r2e_bench_traffic_bf16.json 2> gpurun_out/r2e_bench_traffic_bf16.err; echo "traffic bf16 rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --workload recursive --no-cpu-baseline > gpurun_out/r2e_bench_recursive.json 2> gpurun_out/r2e_bench_recursive.err; echo "recursive rc=$?"; tail -3 gpurun_out/r2e_bench_recursive.err
