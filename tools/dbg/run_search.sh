timeout 120 python profiles/search_bench.py elec 50 2>&1 | tail -3
FLOWTIMES_DFT_TRACE=1 timeout 120 python profiles/search_bench.py elec 5 2>&1 | grep trace
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dft.py tests/test_gpu_r2.py -m gpu -q -x --timeout=300 -k "selector or dft or period or timesblock_forward or golden" 2>&1 | tail -2
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('elec', round(d['ms_per_step'],4), round(d['value']))"
