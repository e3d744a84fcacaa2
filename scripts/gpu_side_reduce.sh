#!/bin/bash
# Runs on the GPU box: the reduction on the side stream (default) against the in-stream reduction (BEAMGPU_NO_SIDE_REDUCE=1).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/side_reduce.log
: > $O
run() { echo "== $*" >> $O; env "$@" >> $O 2>&1; }
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O; tail -3 gpurun_out/pytest_gpu.log >> $O
for cfg in bundle4096 bundle65536; do
  run BEAMGPU_NO_SIDE_REDUCE=1 BEAMGPU_TIMING_DUMP=1 timeout 200 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline
  run BEAMGPU_TIMING_DUMP=1 timeout 200 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline
done
run BEAMGPU_NO_SIDE_REDUCE=1 timeout 200 python bench.py --beams 8192 --steps 20 --warmup 3 --no-cpu-baseline
run timeout 200 python bench.py --beams 8192 --steps 20 --warmup 3 --no-cpu-baseline
grep -v '^{' $O | grep -v "iteration [4-7] of"; grep '^{' $O | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('bench %s beams %d value %.4g e2e %.4g ms/step %.4f frac %.4f share %.4f iters %s parity %s' % (d['config']['config'], d['config']['beams_total'], d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['newton_iterations_per_step'], d['parity']))
"
