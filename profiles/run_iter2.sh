#!/bin/bash
# search-only iteration: DFT tests + trace + microbench + stack bench (no e2e)
timeout 600 python -m pytest tests/test_gpu_dft.py tests/test_gpu_parity.py -m gpu -q --timeout=300 -x 2>&1 | tail -3
FLOWTIMES_DFT_TRACE=1 timeout 120 python profiles/search_bench.py elec 20 2>&1 | tail -5
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value']), 'graph ms', round(d['ms_per_step'],4), 'eager ms', round(d['eager_ms_per_step'],4))"
