"""Build tests/hostbuild/libmpcb_host.so (g++): CPU compilation of the device headers, for tests only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "safe-autonomous-driving-mpc_b200", "csrc")
LIB = os.path.join(HERE, "libmpcb_host.so")


def build(force=False):
    src = os.path.join(HERE, "hostlib.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "mpcb200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-DMPCB_HOST_SOLVER", "-DMPCB_TRACE", "-DMPCB_DEV", "-I", CSRC,
           "-I", os.path.join(ROOT, "include"), "-o", LIB, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
