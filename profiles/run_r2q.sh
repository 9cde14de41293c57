#!/bin/bash
# A/B of the search variants inside the full stack step (graph replay)
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  graph ms', round(d['ms_per_step'],4), 'eager ms', round(d['eager_ms_per_step'],4), d['roofline']['search_kernels'], 'chain', round(d['roofline']['avg_ms'],4))
"; }
run A=1
run FLOWTIMES_NO_TAIL_FOLD=1
run FLOWTIMES_NO_TC_DFT=1
run FLOWTIMES_NO_SEARCH_OVERLAP=1
run FLOWTIMES_NO_SEARCH_OVERLAP=1 FLOWTIMES_NO_TC_DFT=1
run FLOWTIMES_NO_PDL=1
run FLOWTIMES_NO_PDL=1 FLOWTIMES_NO_TC_DFT=1
