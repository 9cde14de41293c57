#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout=300 -k "convs or split" > gpurun_out/r2f_tc.log 2>&1; echo "tc rc=$?"; grep -E "passed|failed|FAILED|rel err" gpurun_out/r2f_tc.log | head
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q --timeout=600 -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --workload etth1 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_etth1.json 2> gpurun_out/r2f_bench_etth1.err; echo "etth1 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_traffic.json 2> gpurun_out/r2f_bench_traffic.err; echo "traffic rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --dtype bf16 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_traffic_bf16.json 2> gpurun_out/r2f_bench_traffic_bf16.err; echo "traffic bf16 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_convs_kernel|tc_gemm_kernel' --launch-skip 18 -c 6 -o gpurun_out/prof_r2f_traffic python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph --workload traffic > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
