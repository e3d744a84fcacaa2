#!/bin/bash
# multi-GPU bench (strong scaling by default): usage gpu_multi2.sh N [extra bench args]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$1; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -5 gpurun_out/bench_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); r=d['roofline']
print('N=$N value %.3f G/s e2e %.3f G/s ms/step %.3f fwd %.3f bwd %.3f share %.3f family %s'%(d['value']/1e9, d['e2e']['value']/1e9 if d['e2e'] else 0, d['ms_per_step'], r['fwd_ms'], r['bwd_ms'], r['kernel_share_of_step'], r['kernel_family'][:24]))
print('parity', d['parity']); print('weak', d['weak']); print('iters', d['newton_iterations'])"
