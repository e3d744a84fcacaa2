#!/bin/bash
# full GPU test-suite, then bench at N=1 (default config) and the two other configs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
make -C oracle > gpurun_out/make.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']
print('N=1 value %.3f G/s e2e %.3f G/s frac %.3f fwd %.3f bwd %.3f share %.3f family %s parity %s'%(d['value']/1e9, d['e2e']['value']/1e9, r['frac'], r['fwd_ms'], r['bwd_ms'], r['kernel_share_of_step'], r['kernel_family'][:20], d['parity']))"
for cfg in bundle4096 long; do
timeout 600 python bench.py --config $cfg > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; tail -3 gpurun_out/bench_$cfg.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$cfg.json')); r=d['roofline']
print('$cfg value %.4f G/s e2e %.4f G/s frac %.3f fwd %.3f bwd %.3f share %.3f family %s cpu %s parity %s'%(d['value']/1e9, d['e2e']['value']/1e9, r['frac'], r['fwd_ms'], r['bwd_ms'], r['kernel_share_of_step'], r['kernel_family'][:20], d['cpu_baseline']['value'], d['parity']))"
done
