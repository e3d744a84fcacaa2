#!/bin/bash
# Runs on the GPU box: the whole GPU test suite, then the three bench configurations (no CPU baseline, short).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
make -C oracle > gpurun_out/make.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -12
for cfg in ${BENCH:-bundle4096 bundle65536 long}; do
python bench.py --config $cfg --no-cpu-baseline ${BENCH_ARGS} 2>gpurun_out/bench_$cfg.err | tee gpurun_out/bench_$cfg.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$cfg value %.3f G/s e2e %s ms/step %.3f kernel_ms %.3f fwd %.3f bwd %.3f share %.3f frac %.3f family %s'%(d['value']/1e9, (d.get('e2e') or {}).get('value'), d['ms_per_step'], r['kernel_ms'], r['fwd_ms'], r['bwd_ms'], r['kernel_share_of_step'], r['frac'], r['kernel_family'][:40]))"
tail -2 gpurun_out/bench_$cfg.err
done
