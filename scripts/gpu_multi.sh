#!/bin/bash
# Runs on a multi-GPU box: bench at N GPUs under torchrun (NCCL reducer inside libbeamgpu).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}; shift
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    if not l.strip().startswith(chr(123)): continue
    d=json.loads(l); print('N',d['n_gpus'],'value %.3f G/s'%(d['value']/1e9),'ms/step',d['ms_per_step'],'e2e',(d['e2e'] or {}).get('value'),'iters',d['newton_iterations_per_step'], d['roofline']['kernel_ms'])
"
