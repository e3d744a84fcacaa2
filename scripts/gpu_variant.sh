#!/bin/bash
# A/B of a compile-time variant on the GPU box: usage gpu_variant.sh "<nvcc extra flags>" <gpu_sizes args...>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
EXTRA="$1"; shift
echo "== baseline"; python -m beamfoam_b200.build --force > /dev/null && python scripts/gpu_sizes.py "$@"
echo "== variant $EXTRA"; BEAMGPU_NVCC_EXTRA="$EXTRA" python -m beamfoam_b200.build --force > /dev/null && python scripts/gpu_sizes.py "$@"
