#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q --timeout=600 > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2i_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench_elec.json 2> gpurun_out/r2i_bench_elec.err; echo "bench rc=$?"; tail -3 gpurun_out/r2i_bench_elec.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --no-cpu-baseline > gpurun_out/r2i_bench_traffic.json 2> gpurun_out/r2i_bench_traffic.err; echo "traffic rc=$?"
