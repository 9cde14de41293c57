#!/bin/bash
# ncu launch list + one --set full capture of the elec stack (r2 state)
bash profiles/ncu_launches.sh r2l
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_conv4_kernel|tc_conv2_kernel|tc_mid_kernel|tc_tail_kernel|tc_gemm2_kernel|spectrum_fft_kernel|spectrum_small_kernel|select_fused_kernel|tc_convs_kernel|tc_gemm_kernel' --launch-skip 18 -c 18 -o gpurun_out/prof_r2l python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/r2l_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
