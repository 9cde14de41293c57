#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_r2.py -m gpu -q --timeout=600 -k "tensor_core_embedding" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2j_pytest.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_e2e_launches.csv python profiles/e2e_launches.py elec > gpurun_out/r2j_e2e.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2j_e2e.log
