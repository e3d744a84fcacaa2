#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests; output tail is what gpurun shows.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
make -C oracle > gpurun_out/make.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -40
