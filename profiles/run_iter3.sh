#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x 2>&1 | tail -4
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${1}_bench.json 2> gpurun_out/${1}_bench.err; echo bench rc=$?; tail -3 gpurun_out/${1}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${1}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'fwd_alone', round(d['e2e']['forward_alone_ms_per_step'],4))
"
timeout 300 python profiles/timeline.py elec e2e 2>&1 | tail -34 | head -24
