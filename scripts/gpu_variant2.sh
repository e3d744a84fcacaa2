cd "${GRAFT_REPO_ROOT:-/root/repo}"
for v in 2 3 4; do echo "== faces min blocks $v"; BEAMGPU_NVCC_EXTRA="-DBEAM_FACES_MIN_BLOCKS=$v" python -m beamfoam_b200.build --force > /dev/null; grep -A3 "beam_update_faces" beamfoam_b200/csrc/ptxas.log | grep -E "registers|spill" | head -2; MODES=sc,1 python scripts/gpu_sizes.py 4096; done
python -m beamfoam_b200.build --force > /dev/null
