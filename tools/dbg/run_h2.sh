export FLOWTIMES_LOG_ERR=gpurun_out/h2_err.log
rm -f gpurun_out/h2_err.log
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout=300 > gpurun_out/h2_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -5 gpurun_out/h2_pytest_tc.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q --timeout=300 > gpurun_out/h2_pytest_par.log 2>&1; echo "pytest parity rc=$?"; tail -8 gpurun_out/h2_pytest_par.log
unset FLOWTIMES_LOG_ERR
for wl in etth1 traffic; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-e2e > gpurun_out/h2_bench_${wl}.json 2> gpurun_out/h2_bench_${wl}.err; echo "$wl h2 rc=$?"
  FLOWTIMES_H2_RING=1 timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-e2e > gpurun_out/h2ring_bench_${wl}.json 2> gpurun_out/h2ring_bench_${wl}.err; echo "$wl h2 ring rc=$?"
  FLOWTIMES_SPLIT_BF16=1 timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-e2e > gpurun_out/s3_bench_${wl}.json 2> gpurun_out/s3_bench_${wl}.err; echo "$wl s3 rc=$?"
done
python - <<'PY'
import json
for t in ["h2","h2ring","s3"]:
    for wl in ["etth1","traffic"]:
        try:
            d=json.load(open(f"gpurun_out/{t}_bench_{wl}.json"))
            ch={k["kernel"]:round(k["avg_ms"]*1e3,1) for k in d["roofline"]["chain_kernels"]}
            print(t, wl, round(d["ms_per_step"],3), ch)
        except Exception as e:
            print(t, wl, "failed", e)
PY
