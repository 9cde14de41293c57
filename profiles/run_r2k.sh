#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q --timeout=300 > gpurun_out/r2k_bw.log 2>&1; echo "bw rc=$?"; grep -E "passed|failed|Error|assert" gpurun_out/r2k_bw.log | head -20
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_tc.py -m gpu -q --timeout=600 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
