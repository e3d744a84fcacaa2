#!/bin/bash
# Runs on the GPU box: launch list + one full ncu capture of the two launches of a Newton iteration.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline $@"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:beam_(forward|backward)" -s 12 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full.log
