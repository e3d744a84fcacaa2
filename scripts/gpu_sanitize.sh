#!/bin/bash
# Runs on the GPU box: compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over the kernel families on small bundles.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
make -C oracle > gpurun_out/make.log 2>&1
SEL="tests/test_gpu_coop.py::test_newmark_bundle tests/test_gpu_coop.py::test_ragged_bundle tests/test_gpu_golden_and_edges.py::test_step_outputs_and_async_bc_upload"
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -m gpu -q -x > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done
