#!/bin/bash
# one build -> measure iteration: GPU tests, search microbench, stack bench, kernel timeline of the replayed graph
tag=${1:-iter}
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x 2>&1 | tail -4
timeout 120 python profiles/search_bench.py elec 50
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo bench rc=$?; tail -3 gpurun_out/${tag}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'fwd_alone', round(d['e2e']['forward_alone_ms_per_step'],4), d['roofline']['search_kernels'])
"
timeout 300 python profiles/timeline.py elec 2>&1 | tail -16
