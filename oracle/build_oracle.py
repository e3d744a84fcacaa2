"""Builds oracle/_build/libbeamoracle.so (gcc).  TEST INFRASTRUCTURE ONLY — used by tests/,
__graft_entry__ and bench.py's cpu_baseline / --impl reference legs, never by beamfoam_b200."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libbeamoracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, "beam_oracle.c"), os.path.join(HERE, "..", "include", "beamgpu.h")]
    stale = not os.path.exists(ORACLE_SO) or any(os.path.getmtime(ORACLE_SO) < os.path.getmtime(s) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-B", "-C", HERE], check=True, capture_output=True)
    return ORACLE_SO


if __name__ == "__main__":
    print(build(force=True))
