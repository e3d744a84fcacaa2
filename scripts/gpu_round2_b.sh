#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
make -C oracle > gpurun_out/make.log 2>&1
timeout 900 python -m pytest tests/test_gpu_coop.py -m gpu -q -x 2>&1 | tee gpurun_out/pytest_coop.log | tail -8
timeout 600 python scripts/gpu_sizes.py ${@:-4096 8192 16384} 2>&1 | tee gpurun_out/sizes.log
