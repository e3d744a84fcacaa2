#!/bin/bash
# Runs on the GPU box: parity of the kernel families (tests/test_gpu_coop.py), then per-iteration timings by family and size.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
make -C oracle > gpurun_out/make.log 2>&1
timeout 900 python -m pytest tests/test_gpu_coop.py -m gpu -q -x 2>&1 | tee gpurun_out/pytest_coop.log | tail -15
MODES=${MODES:-s,sc,w16,w32} timeout 600 python scripts/gpu_sizes.py ${SIZES:-4096 8192 16384} 2>&1 | tee gpurun_out/sizes_ws.log
if [ -n "$BENCH" ]; then
for cfg in $BENCH; do
python bench.py --config $cfg --no-cpu-baseline --no-e2e 2>gpurun_out/bench_$cfg.err | tee gpurun_out/bench_$cfg.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$cfg value %.3f G/s ms/step %.3f kernel_ms %.3f fwd %.3f bwd %.3f share %.3f family %s'%(d['value']/1e9, d['ms_per_step'], r['kernel_ms'], r['fwd_ms'], r['bwd_ms'], r['kernel_share_of_step'], r['kernel_family']))"
done
fi
