set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r1G_bench_elec.json 2> gpurun_out/bench_elec.err; echo "bench rc=$?"
cat gpurun_out/r1G_bench_elec.json
timeout 300 python bench.py --steps 10 --warmup 3 --workload etth1 > gpurun_out/r1G_bench_etth1.json 2> gpurun_out/bench_etth1.err; echo "etth1 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --workload traffic --no-cpu-baseline > gpurun_out/r1G_bench_traffic.json 2> gpurun_out/bench_traffic.err; echo "traffic rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1G_bench_ref.json 2>&1; echo "ref rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 bash profiles/ncu_launches.sh r1G
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_conv4_kernel|tc_conv2_kernel|tc_mid_kernel|tc_tail_kernel|tc_gemm2_kernel|spectrum_fft_kernel|select_fused_kernel' --launch-skip 18 -c 18 -o gpurun_out/prof_r1G python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
