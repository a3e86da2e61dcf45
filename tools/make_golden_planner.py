"""Generate planner golden vectors from the UNMODIFIED reference (imported from /root/reference).

Build-container only.  The reference's trajectory_planning.py imports path_planning, which cannot be imported here
(pymap3d absent, API key required at import, invalid annotation on Python 3.12 -- SURVEY.md C6), so an empty stub
module named ``path_planning`` is placed in sys.modules first; TrajectoryOptimizer itself is used unmodified.
k_ref_fun / v_max_fun are the synthetic substitutes of SURVEY.md 8(d): TrajectoryLoader.interp_k, constant v_max.

Writes tests/golden/planner_traj{1,2,3}.npz:  windows of N=12 intervals cut from the committed trajectories (half of
them perturbed), with the value of every constraint closure and the cost as returned by the reference.

    python tools/make_golden_planner.py
"""
import os
import sys
import types

import numpy as np
import scipy

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

stub = types.ModuleType("path_planning")
stub.get_route = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub"))
stub.get_path_and_speed_limits = stub.get_route
sys.modules["path_planning"] = stub
sys.path.insert(0, REF)

import trajectory_planning as tp          # noqa: E402
from trajectory_loader import TrajectoryLoader   # noqa: E402

N = 12
DT = 0.3


def main():
    for i in (1, 2, 3):
        L = TrajectoryLoader(os.path.join(REF, "trajectories", f"trajectory{i}.json"))
        import json
        with open(os.path.join(REF, "trajectories", f"trajectory{i}.json")) as f:
            S_all = np.array(json.load(f)["S"], dtype=np.float64)
        X_all, U_all = L.X_ref, L.U_ref
        K = len(X_all)
        v_max = float(X_all[:, 4].max())
        s_total = float(X_all[-1, 0])
        k_ref_fun = lambda s: float(L.interp_k(s))            # noqa: E731
        v_min_fun = lambda s: 0                               # noqa: E731  (trajectory_planning.py:476-477)
        v_max_fun = lambda s: v_max                           # noqa: E731
        opt = tp.TrajectoryOptimizer(horizon=N * DT, N=N, dt=DT)
        rng = np.random.default_rng(500 + i)
        starts = list(range(0, K - 1 - N, max(1, (K - 1 - N) // 24)))[:24] + [K - 1 - N]
        zs, x0s, fin, defs, nodes, ctrls, costs, init, term, s_tg = [], [], [], [], [], [], [], [], [], []
        for w, a in enumerate(starts):
            X = X_all[a:a + N + 1].copy()
            U = U_all[a:a + N].copy()
            S = S_all[a:a + N].copy()
            if w % 2 == 1:                                     # perturbed: off-knot s, non-zero slack and defects
                X += rng.normal(0, 1, X.shape) * np.array([0.3, 0.05, 0.02, 0.01, 0.3])
                U += rng.normal(0, 1, U.shape) * np.array([0.05, 0.3])
                S = np.abs(rng.normal(0, 0.05, S.shape))
            z = opt.pack(X, U, S)
            x0 = X_all[a].copy()
            is_final = (a == K - 1 - N)
            s_target = float(X_all[a + N, 0])
            cons = opt.constraints(x0, s_target, k_ref_fun, v_min_fun, v_max_fun, is_final)
            vals = [np.atleast_1d(np.asarray(c["fun"](z), dtype=np.float64)) for c in cons]
            p = 0
            d = np.array(vals[p:p + N]); p += N
            ini = vals[p]; p += 1
            nterm = 2 if is_final else 1
            te = np.concatenate(vals[p:p + nterm] + ([np.array([np.nan])] if nterm == 1 else [])); p += nterm
            nd = np.zeros((N + 1, 6))
            for k in range(N + 1):
                nd[k, 0:4] = [vals[p][0], vals[p + 1][0], vals[p + 2][0], vals[p + 3][0]]; p += 4
            for k in range(N + 1):
                nd[k, 4:6] = [vals[p][0], vals[p + 1][0]]; p += 2
            ct = np.zeros((N, 5))
            for k in range(N):
                ct[k] = [vals[p + q][0] for q in range(5)]; p += 5
            assert p == len(vals)
            zs.append(z); x0s.append(x0); fin.append(is_final); defs.append(d); nodes.append(nd); ctrls.append(ct)
            costs.append(float(opt.cost(z, x0, s_total))); init.append(ini); term.append(te); s_tg.append(s_target)
        np.savez_compressed(os.path.join(OUT, f"planner_traj{i}.npz"), z=np.array(zs), x0=np.array(x0s),
                            is_final=np.array(fin), defect=np.array(defs), node_rows=np.array(nodes),
                            ctrl_rows=np.array(ctrls), cost=np.array(costs), initial=np.array(init),
                            terminal=np.array(term), s_target=np.array(s_tg), s_total=s_total, v_max=v_max,
                            N=N, dt=DT, starts=np.array(starts),
                            versions=np.array([np.__version__, scipy.__version__]))
        print(i, "windows", len(zs), "max |defect| (committed sign)", float(np.abs(np.array(defs)).max()))


if __name__ == "__main__" and "--chunk" not in sys.argv:
    main()


def make_chunk_solution():
    """One chunk solved by the UNMODIFIED reference ``TrajectoryOptimizer.optimize`` (SLSQP + finite differences,
    :351-390) with the synthetic route functions: the known answer for the planner driver."""
    import time
    L = TrajectoryLoader(os.path.join(REF, "trajectories", "trajectory1.json"))
    v_max = float(L.X_ref[:, 4].max())
    s_total = float(L.X_ref[-1, 0])
    opt = tp.TrajectoryOptimizer(horizon=N * DT, N=N, dt=DT)
    x0 = np.array([0.0, 0.0, 0.0, 0.0, 0.0])
    t0 = time.time()
    X, U, S = opt.optimize(x0, 20.0, s_total, lambda s: float(L.interp_k(s)), lambda s: 0, lambda s: v_max, False)
    z = opt.pack(X, U, S)
    np.savez_compressed(os.path.join(OUT, "planner_chunk_traj1.npz"), X=X, U=U, S=S, cost=float(opt.cost(z, x0, s_total)),
                        x0=x0, s_target=20.0, s_total=s_total, v_max=v_max, N=N, dt=DT, seconds=time.time() - t0,
                        versions=np.array([np.__version__, scipy.__version__]))
    print("chunk: cost", float(opt.cost(z, x0, s_total)), "s_N", X[-1, 0], "seconds", time.time() - t0)


if __name__ == "__main__" and "--chunk" in sys.argv:
    make_chunk_solution()
