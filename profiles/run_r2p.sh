#!/bin/bash
# tensor-core DFT: tests, microbench of the search alone, ncu timing + full capture of the search kernels
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x 2>&1 | tail -8
timeout 120 python profiles/search_bench.py elec 50
timeout 120 python profiles/search_bench.py etth1 50
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2p_search_launches.csv python profiles/search_bench.py elec 1 > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_dft_kernel' --launch-skip 20 -c 2 -o gpurun_out/prof_r2p python profiles/search_bench.py elec 1 > gpurun_out/r2p_ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import csv,collections
agg=collections.OrderedDict()
lines=[l for l in open('gpurun_out/r2p_search_launches.csv') if not l.startswith('==')]
for row in csv.DictReader(lines):
    n=row['Kernel Name'][:60]; v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u.startswith('n') else v
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print(f"{k:62s} {n:4d} {t/n:8.2f} us")
PY
FLOWTIMES_NO_TAIL_FOLD=1 timeout 120 python profiles/search_bench.py elec 50
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2p_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['forward_alone_ms_per_step'], d['roofline']['search_kernels'])
"
