#!/bin/bash
# round 2, evidence pass after the fp16-pair fp32 route, the granule layout, stacked conv4 units and quad tail items (one GPU): tests, smoke, bench lines of every workload, reference arm, ncu launch list,
# ncu --set full capture of one stack forward, CUPTI timelines of the replayed graphs
t=r4
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_conv4_kernel|tc_mid_kernel|tc_tail_kernel|tc_gemm2_kernel|tc_dft_kernel' --launch-skip 24 -c 12 -o gpurun_out/prof_${t} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/${t}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 200 python profiles/timeline.py elec > /dev/null 2>&1; cp gpurun_out/timeline_elec.txt gpurun_out/${t}_timeline_elec.txt
timeout 200 python profiles/timeline.py elec e2e > /dev/null 2>&1; cp gpurun_out/timeline_elec_e2e.txt gpurun_out/${t}_timeline_elec_e2e.txt
timeout 120 python profiles/search_bench.py elec 50 > gpurun_out/${t}_search_bench.txt 2>&1
timeout 120 python profiles/search_bench.py etth1 50 >> gpurun_out/${t}_search_bench.txt 2>&1
FLOWTIMES_DFT_TRACE=1 timeout 120 python profiles/search_bench.py elec 5 2>&1 | grep trace >> gpurun_out/${t}_search_bench.txt
ls -la gpurun_out/*.ncu-rep
FLOWTIMES_LOG_ERR=gpurun_out/${t}_parity_margins.txt timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q --timeout=300 > /dev/null 2>&1; sort -r gpurun_out/${t}_parity_margins.txt | head -60 > gpurun_out/${t}_parity_margins_top.txt
