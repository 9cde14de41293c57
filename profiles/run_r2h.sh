#!/bin/bash
set -x
rm -f gpurun_out/r2h_errors.txt
FLOWTIMES_LOG_ERR=gpurun_out/r2h_errors.txt timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q --timeout=600 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
