"""Development aid: run the solver source compiled for the host (tests/hostbuild) on the golden problem sets and on
a slice of the Monte-Carlo set; report parity against the converged reference answers and iteration statistics."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "hostbuild"))
import build as hb   # noqa: E402
from oracle import tracker_port as P   # noqa: E402

lib = C.CDLL(hb.build())
vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731


def host_solve(tab, x0, obs, n, params=None):
    global LAST_PASS
    B = len(n)
    s = np.ascontiguousarray(tab.s)
    y = np.ascontiguousarray(tab.X[:, 1:5])
    u = np.ascontiguousarray(tab.U[: tab.Ku])
    last4 = np.ascontiguousarray(tab.X[-1, 1:5])
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    n = np.ascontiguousarray(n, dtype=np.int32)
    U = np.zeros((B, 10)); st = np.zeros(B, np.int32); it = np.zeros((B, 2), np.int32); obj = np.zeros(B)
    LAST_PASS = np.zeros(B, np.int32)
    rc = lib.host_solve_batch(vp(s), vp(y), vp(u), C.c_int(tab.K), C.c_int(tab.Ku), C.c_double(tab.s_max), vp(last4),
                              params, C.c_int(B), vp(x0), vp(obs), vp(n), vp(U), vp(st), vp(it), vp(obj), vp(LAST_PASS))
    assert rc == 0
    return U, st, it, obj


def params_from_env():
    sys.path.insert(0, ROOT)
    import importlib
    L = importlib.import_module("safe_autonomous_driving_mpc_b200._lib")
    p = L.Params()
    L.load().mpcb_default_params(C.byref(p))
    for k, v in os.environ.items():
        if k.startswith("P_"):
            f = k[2:]
            setattr(p, f, type(getattr(p, f))(float(v)) if not isinstance(getattr(p, f), int) else int(v))
    return p


def main():
    global PARAMS
    PARAMS = params_from_env()
    for name, i in (("solve_traj1", 1), ("solve_traj2", 2), ("solve_traj3", 3), ("solve_mc_traj3", 3)):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        tab = P.RefTable.from_npz(os.path.join(ROOT, "data", f"trajectory{i}.npz"))
        t0 = time.time()
        U, st, it, obj = host_solve(tab, g["x0"], g["obs_sv"], g["n_obs"], C.byref(PARAMS))
        dt = time.time() - t0
        pin = g["pinned"]
        err = np.abs(U - g["U_conv"]).max(axis=1)
        jerr = np.abs(obj - g["J_conv"]) / np.maximum(np.abs(g["J_conv"]), 1.0)
        print(f"{name}: pinned {pin.sum()}/{len(pin)} err max {err[pin].max():.2e} n>1e-4 {(err[pin] > 1e-4).sum()} "
              f"Jrel {jerr[pin].max():.1e} | status pinned {np.bincount(st[pin], minlength=3)} unpinned "
              f"{np.bincount(st[~pin], minlength=3)} | rounds {it[:, 0].mean():.2f} iters {it[:, 1].mean():.0f}/"
              f"{it[:, 1].max()} | pass2 {(LAST_PASS == 2).sum()} | {dt * 1e3 / len(pin):.2f} ms/solve")
    tab = P.RefTable.from_npz(os.path.join(ROOT, "data", "trajectory3.npz"))
    x0, obs, n = P.monte_carlo_problems(tab, 65536)
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    t0 = time.time()
    U, st, it, obj = host_solve(tab, x0[:m], obs[:m], n[:m], C.byref(PARAMS))
    dt = time.time() - t0
    print(f"MC first {m}: status {np.bincount(st, minlength=3)} rounds {it[:, 0].mean():.2f} iters mean "
          f"{it[:, 1].mean():.1f} max {it[:, 1].max()} | pass2 {(LAST_PASS == 2).sum()} | {dt * 1e3 / m:.3f} ms/solve")
    np.savez("/tmp/host_mc.npz", U=U, st=st, it=it, obj=obj)


if __name__ == "__main__":
    main()
