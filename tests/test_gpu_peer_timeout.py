"""A peer that never publishes its norm sums must not hang the GPU: two processes share cuda:0 and attach each other's
exchange buffers (bf_comm_ipc_export / bf_comm_ipc_attach); rank 1 never calls evolve().  Rank 0's kernels give the peer up
after BEAMGPU_PEER_TIMEOUT_S and bf_evolve returns BF_ERR_COMM instead of spinning on the device for ever."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import ctypes as C, os, sys, time
    sys.path.insert(0, %r)
    from beamfoam_b200 import _lib as L, cases, coupledTotalLagNewtonRaphsonBeam, load_library
    rank, d = int(sys.argv[1]), sys.argv[2]
    lib = load_library()
    m = coupledTotalLagNewtonRaphsonBeam(cases.multiBeamDensity(64, 20, jitter=True), library=lib)
    h = (C.c_char * 64)()
    assert lib.bf_comm_ipc_export(m._ctx, h) == 0
    open(os.path.join(d, "h%%d.tmp" %% rank), "wb").write(bytes(h)); os.rename(os.path.join(d, "h%%d.tmp" %% rank), os.path.join(d, "h%%d" %% rank))
    other = os.path.join(d, "h%%d" %% (1 - rank))
    while not os.path.exists(other):
        time.sleep(0.01)
    hs = [None, None]; hs[rank] = bytes(h); hs[1 - rank] = open(other, "rb").read()
    assert lib.bf_comm_ipc_attach(m._ctx, 2, rank, b"".join(hs)) == 0, lib.bf_last_error(m._ctx)
    open(os.path.join(d, "attached%%d" %% rank), "w").close()
    while not os.path.exists(os.path.join(d, "attached%%d" %% (1 - rank))):
        time.sleep(0.01)
    if rank == 0:
        m.loop()
        m.updateCoeffs(m.time)
        assert lib.bf_begin_step(m._ctx, float(m.case.deltaT)) == 0
        out = L.bf_step_out()
        t0 = time.time()
        rc = lib.bf_evolve(m._ctx, C.byref(m.tol), C.byref(out))
        print("RC", rc, L.BF_ERR_COMM, "%%.2f" %% (time.time() - t0), lib.bf_last_error(m._ctx).decode(), flush=True)
        open(os.path.join(d, "done"), "w").close()
    else:
        while not os.path.exists(os.path.join(d, "done")):  # alive, attached, but never publishing
            time.sleep(0.05)
""") % ROOT


def test_lost_peer_is_an_error_not_a_hang(gpu_lib, tmp_path):
    env = dict(os.environ, BEAMGPU_PEER_TIMEOUT_S="1.0")
    procs = [subprocess.Popen([sys.executable, "-c", WORKER, str(r), str(tmp_path)], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in (0, 1)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert procs[0].returncode == 0, outs[0][1][-2000:]
    assert procs[1].returncode == 0, outs[1][1][-2000:]
    line = [l for l in outs[0][0].splitlines() if l.startswith("RC")][0].split(None, 4)
    assert int(line[1]) == int(line[2]), line          # BF_ERR_COMM
    assert float(line[3]) < 30.0, line                 # gave up after the timeout (a few waits of 1 s), did not hang
    assert "did not arrive" in line[4]
