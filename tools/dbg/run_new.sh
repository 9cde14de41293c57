timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/new_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/new_pytest.log
