#!/bin/bash
# Round-2 profiles (runs on the GPU box): launch list of the default bench command, then one ncu --set full capture of the
# Newton-iteration kernels of every kernel family at the bundle size it serves.  Every command first runs WITHOUT ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BEAMGPU_NO_PIPELINE=1   # one host round trip per iteration: no speculative (early-exit) launches among the captured ones
WHAT=${@:-launches split65536 perbeam32768 splitcell8192 splitgroup4096 group4096}   # gpurun copies back at most 64 MiB: two captures per call
want() { [[ " $WHAT " == *" $1 "* ]]; }
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-verify"
if want launches; then
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
fi
cap() { # name modes beams kernel-regex count.  Skips 5 iterations: the captured one is the SECOND of a time step (the first one also applies the Newmark predictor and moves 12 more doubles per cell)
  want $1 || return 0
  MODES=$2 python scripts/gpu_sizes.py $3 > gpurun_out/plain_$1.log 2>&1 &&
  MODES=$2 ncu --set full --clock-control none --import-source on -k "regex:$4" -s $((5 * $5)) -c $5 -f -o gpurun_out/prof_$1 python scripts/gpu_sizes.py $3 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"; cat gpurun_out/plain_$1.log
}
cap split65536 s 65536 '^beam_(forward|backward)_split$' 2
cap perbeam32768 0 32768 '^beam_(forward|backward)$' 2
cap splitcell8192 sc 8192 '^beam_(forward_split|backsub_split|update)$' 3
cap splitgroup4096 w16 4096 '^beam_(forward_ws|backsub_split|update)$' 3
cap group4096 1 4096 '^beam_(forward_ws|backsub|update)$' 3
