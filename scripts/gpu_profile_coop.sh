#!/bin/bash
# ncu --set full of the cooperative kernels at a sub-wave bundle size (default 4096 beams x 200 cells)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NB=${1:-4096}
export MODES=1
python scripts/gpu_sizes.py $NB > gpurun_out/plain_coop.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:beam_(forward_ws|backsub|update_faces|update_cells)" -s 16 -c 4 -f -o gpurun_out/prof_coop python scripts/gpu_sizes.py $NB > gpurun_out/ncu_coop.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_coop.log; tail -2 gpurun_out/ncu_coop.log
