"""Golden verdicts of the UNMODIFIED reference's trajectory_tracking_check (sanity_checks.py:79-184) on recorded drives
that PASS and on synthetic variants that FAIL one or several items.  Build-container only (imports /root/reference).

Base histories are the reference's own closed-loop logs (tests/golden/closed_loop_traj{2,3}.npz); each case edits one
thing (truncate the drive, push d past 1.5 m, a control past its limit +- 0.1, move the car to within 1 m, turn the
light red at the moment of passing, ...).  The reference function prints one line per item and returns the overall
verdict; both are stored.  tests/test_gpu_sim.py feeds the same histories to mpcb_check_histories.

    python tools/make_golden_sanity.py
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import scipy

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
import sanity_checks as SC   # noqa: E402
import trajectory_tracking as tt   # noqa: E402

ITEMS = [("destination", "Destination reached"), ("on_road", "Stayed on road"),
         ("steering", "Steering controls within limits"), ("acceleration", "Acceleration controls within limits"),
         ("obstacle", "Dynamic Obstacle Avoided"), ("light", "Traffic Light Respected")]


def run_reference(case):
    fsm = types.SimpleNamespace(dynamic_obstacle=bool(case["dyn"]), traffic_light=bool(case["tl"]), tl_pos=case["tl_pos"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        passed = SC.trajectory_tracking_check(tt.TrajectoryTracker(), list(case["hx"]), list(case["hu"]),
                                              list(np.zeros(len(case["hu"]))), list(case["obs"]), list(case["tls"]), fsm,
                                              case["s_total"])
    out = buf.getvalue()
    items = {}
    for key, text in ITEMS:
        line = [ln for ln in out.splitlines() if ln.startswith(text)]
        items[key] = (": True" in line[0]) if line else True      # item not evaluated (no car / no light): nothing failed
    return bool(passed), items, out


def base(i):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"closed_loop_traj{i}.npz"))
    tl_pos = {2: 550.0, 3: 2000.0}[i]
    return dict(hx=z["hist_x"].copy(), hu=z["hist_u"].copy(), obs=np.asarray(z["hist_obs_s"], float).copy(),
                tls=[str(t) for t in z["hist_tl"]], dyn=1, tl=1, tl_pos=tl_pos, s_total=float(z["hist_x"][-1, 0] + 0.5), traj=i)


def main():
    cases = []

    def add(name, c):
        c["name"] = name
        cases.append(c)

    for i in (2, 3):
        add(f"traj{i} as recorded", base(i))
    c = base(2); n = int(0.6 * len(c["hu"]))
    c["hx"], c["hu"], c["obs"], c["tls"] = c["hx"][: n + 1], c["hu"][:n], c["obs"][:n], c["tls"][:n]
    add("stopped short of the destination", c)
    c = base(2); c["hx"][300, 1] = 1.6; add("left the road (d = 1.6 m)", c)
    c = base(2); c["hx"][300, 1] = -1.5; add("d = -1.5 m exactly (still on the road)", c)
    c = base(2); c["hx"][-1, 1] = 1.7; add("left the road in the FINAL state only", c)
    c = base(2); c["hu"][100, 0] = 0.75; add("steering above u1_max + 0.1", c)
    c = base(2); c["hu"][100, 0] = -0.71; add("steering below u1_min - 0.1", c)
    c = base(2); c["hu"][100, 0] = 0.6 + 0.1; add("steering exactly u1_max + 0.1 (allowed)", c)
    c = base(2); c["hu"][50, 1] = 4.2; add("acceleration above u2_max + 0.1", c)
    c = base(2); c["hu"][50, 1] = -5.11; add("braking below u2_min - 0.1", c)
    c = base(2)
    m = ~np.isnan(c["obs"]); t = np.where(m)[0][40]
    c["obs"][t] = c["hx"][t, 0] + 0.5; add("car within 0.5 m", c)
    c = base(2); c["obs"][t] = c["hx"][t, 0] + 1.0; add("car at exactly 1.0 m (allowed)", c)
    c = base(2); c["obs"][t] = c["hx"][t, 0] - 3.0; add("car BEHIND the vehicle (negative gap)", c)
    c = base(2); c["obs"][t] = c["hx"][t, 0] + 0.5; c["dyn"] = 0; add("car within 0.5 m but no dynamic obstacle configured", c)
    c = base(2); k = int(np.where(c["hx"][:, 0] > c["tl_pos"])[0][0]); c["tls"][k] = "RED"; add("ran the red light", c)
    c = base(2); c["tls"][k] = "RED"; c["tl"] = 0; add("red at the moment of passing but no light configured", c)
    c = base(2); c["tls"][k - 1] = "RED"; c["tls"][k] = "GREEN"; add("red one step before passing only", c)
    c = base(3); c["hx"][900, 1] = -2.0; c["hu"][10, 1] = 4.5; c["obs"][np.where(~np.isnan(c["obs"]))[0][5]] = c["hx"][np.where(~np.isnan(c["obs"]))[0][5], 0] + 0.2
    add("three items at once (road, acceleration, car)", c)
    c = base(3); k3 = int(np.where(c["hx"][:, 0] > c["tl_pos"])[0][0]); c["tls"][k3] = "RED"; c["hu"][5, 0] = 0.9
    add("trajectory3: red light and steering", c)

    B = len(cases)
    T = max(len(c["hu"]) for c in cases)
    hist_x = np.full((T, B, 5), np.nan); hist_u = np.full((T, B, 2), np.nan); hist_obs = np.full((T, B), np.nan)
    hist_tl = np.full((T, B), -1, np.int32)
    x_final = np.zeros((B, 5)); steps = np.zeros(B, np.int32)
    dyn = np.zeros(B, np.int32); tl = np.zeros(B, np.int32); tl_pos = np.zeros(B); s_total = np.zeros(B)
    passed = np.zeros(B, bool); items = np.zeros((B, len(ITEMS)), bool); names = []; stdout = []
    for b, c in enumerate(cases):
        p, it, out = run_reference(c)
        n = len(c["hu"])
        hist_x[:n, b] = c["hx"][:-1]; hist_u[:n, b] = c["hu"]; hist_obs[:n, b] = c["obs"]
        hist_tl[:n, b] = [1 if s == "GREEN" else 0 for s in c["tls"]]
        x_final[b] = c["hx"][-1]; steps[b] = n
        dyn[b], tl[b], tl_pos[b], s_total[b] = c["dyn"], c["tl"], c["tl_pos"], c["s_total"]
        passed[b] = p; items[b] = [it[k] for k, _ in ITEMS]; names.append(c["name"]); stdout.append(out)
        print(f"{b:2d} {c['name']:55s} passed={p} " + " ".join(f"{k}={int(it[k])}" for k, _ in ITEMS))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sanity_cases.npz"),
                        versions=np.array([np.__version__, scipy.__version__]), names=np.array(names),
                        item_names=np.array([k for k, _ in ITEMS]), hist_x=hist_x, hist_u=hist_u, hist_obs=hist_obs,
                        hist_tl=hist_tl, x_final=x_final, steps=steps, dynamic_obstacle=dyn, traffic_light=tl,
                        tl_pos=tl_pos, s_total=s_total, ref_passed=passed, ref_items=items, ref_stdout=np.array(stdout))


if __name__ == "__main__":
    main()
