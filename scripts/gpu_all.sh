#!/bin/bash
# Runs on the GPU box: parity tests, then the bench (args are passed to bench.py).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
make -C oracle > gpurun_out/make.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | grep -E "^E  .*(Error|assert [0-9])|FAILED|passed|failed" | head -30
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
r=d['roofline']
print(f"value={d['value']/1e9:.3f} G/s  ms/step={d['ms_per_step']:.2f} iters/step={d['newton_iterations_per_step']}  kernel_ms={r['kernel_ms']:.3f} frac={r['frac']:.4f}  e2e={(d['e2e'] or {}).get('value',0)/1e9:.3f} G/s  cpu={(d['cpu_baseline'] or {}).get('value',0)/1e6:.2f} M/s clocks={d['clocks']}")
PY
