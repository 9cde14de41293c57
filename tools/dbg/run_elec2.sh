for i in 1 2; do
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ab_elec_$i.json 2> gpurun_out/ab_elec_$i.err; echo "elec rc=$?"
done
python - <<'PY'
import json
for i in (1,2):
    d=json.load(open(f"gpurun_out/ab_elec_{i}.json"))
    ch={k["kernel"][:8]:round(k["avg_ms"]*1e3,1) for k in d["roofline"]["chain_kernels"]}
    print(i, round(d["ms_per_step"],4), round(d["value"]), ch)
PY
