#!/bin/bash
# round 2, FINAL evidence pass (one GPU): tests, smoke, bench lines of every workload, reference arm, launch list, timelines, search trace
t=r7
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${t}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${t}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${t}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${t}_smoke.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/${t}_bench_elec.json 2> gpurun_out/${t}_bench_elec.err; echo "bench rc=$?"; tail -2 gpurun_out/${t}_bench_elec.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${t}_bench_ref.json 2> gpurun_out/${t}_bench_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/${t}_bench_ref.err
timeout 300 python bench.py --steps 3 --warmup 3 --workload recursive --no-cpu-baseline > gpurun_out/${t}_bench_recursive.json 2> gpurun_out/${t}_bench_recursive.err; echo "recursive rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --workload etth1 --no-cpu-baseline > gpurun_out/${t}_bench_etth1.json 2> gpurun_out/${t}_bench_etth1.err; echo "etth1 rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --workload etth1 --dtype bf16 --no-cpu-baseline > gpurun_out/${t}_bench_etth1_bf16.json 2> gpurun_out/${t}_bench_etth1_bf16.err; echo "etth1 bf16 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --no-cpu-baseline > gpurun_out/${t}_bench_traffic.json 2> gpurun_out/${t}_bench_traffic.err; echo "traffic rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --dtype bf16 --no-cpu-baseline > gpurun_out/${t}_bench_traffic_bf16.json 2> gpurun_out/${t}_bench_traffic_bf16.err; echo "traffic bf16 rc=$?"
bash profiles/ncu_launches.sh ${t}
timeout 200 python profiles/timeline.py elec > /dev/null 2>&1; cp gpurun_out/timeline_elec.txt gpurun_out/${t}_timeline_elec.txt
timeout 200 python profiles/timeline.py elec e2e > /dev/null 2>&1; cp gpurun_out/timeline_elec_e2e.txt gpurun_out/${t}_timeline_elec_e2e.txt
timeout 120 python profiles/search_bench.py elec 50 > gpurun_out/${t}_search_bench.txt 2>&1
timeout 120 python profiles/search_bench.py etth1 50 >> gpurun_out/${t}_search_bench.txt 2>&1
FLOWTIMES_DFT_TRACE=1 timeout 120 python profiles/search_bench.py elec 5 2>&1 | grep trace >> gpurun_out/${t}_search_bench.txt
