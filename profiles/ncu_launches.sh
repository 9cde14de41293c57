#!/bin/bash
# usage (on the GPU box, from the repo root): profiles/ncu_launches.sh <tag> [bench args...]
# 1) plain run must exit 0, 2) ncu timing pass of the same command -> gpurun_out/<tag>_launches.csv
tag=$1; shift
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
