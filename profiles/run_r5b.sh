#!/bin/bash
# ncu --set full captures of one stack forward of the etth1 (fp32: fp16-pair GEMMs, tc_convs row mode) and recursive
# (bf16: stacked tc_conv4 units, granule tc_mid, four-window tc_tail items) workloads; each after a plain run exited 0
t=r5b
timeout 300 python bench.py --steps 1 --warmup 1 --workload etth1 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/${t}_plain_etth1.log 2>&1 || { echo "plain etth1 failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel|tc_convs_kernel|split_h2_kernel|aggregate|spectrum_fft_kernel|select_fused' --launch-skip 40 -c 14 -o gpurun_out/prof_${t}_etth1 python bench.py --steps 1 --warmup 1 --workload etth1 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/${t}_ncu_etth1.log 2>&1; echo "ncu etth1 rc=$?"
timeout 300 python bench.py --steps 1 --warmup 1 --workload recursive --batch 4096 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/${t}_plain_rec.log 2>&1 || { echo "plain recursive failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_conv4_kernel|tc_mid_kernel|tc_tail_kernel|tc_gemm2_kernel|spectrum_small_kernel|select_fused' --launch-skip 60 -c 14 -o gpurun_out/prof_${t}_rec python bench.py --steps 1 --warmup 1 --workload recursive --batch 4096 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/${t}_ncu_rec.log 2>&1; echo "ncu recursive rc=$?"
ls -la gpurun_out/prof_${t}_*.ncu-rep
