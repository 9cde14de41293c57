#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout=300 -k "convs or split" > gpurun_out/r2c_tc.log 2>&1; echo "tc rc=$?"; grep -E "passed|failed|FAILED|rel err" gpurun_out/r2c_tc.log | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_convs_kernel|tc_gemm_kernel' --launch-skip 12 -c 6 -o gpurun_out/prof_r2c_etth1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph --workload etth1 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
