#!/bin/bash
tag=${1:-iter}
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x 2>&1 | tail -4
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null > gpurun_out/${tag}_bench.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'graph ms', round(d['ms_per_step'],4), 'eager ms', round(d['eager_ms_per_step'],4))
print([ (k['kernel'][:12], round(k['avg_ms']*1e3,1)) for k in d['roofline']['chain_kernels']])"
timeout 300 python profiles/timeline.py elec 2>&1 | tail -9
