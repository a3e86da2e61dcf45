"""Pack the reference's committed planner outputs (trajectories/trajectory{1,2,3}.json: keys X (K,5),
U (K-1,2), S (K-1,)) into binary fixtures data/trajectory{i}.npz, bit-exact float64.

Run in the build container only (needs /root/reference):  python tools/import_trajectories.py
The .npz files are committed; nothing at test/bench time reads /root/reference.
"""
import json
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")
os.makedirs(OUT, exist_ok=True)
for i in (1, 2, 3):
    with open(os.path.join(REF, "trajectories", f"trajectory{i}.json")) as f:
        d = json.load(f)
    X = np.array(d["X"], dtype=np.float64)
    U = np.array(d["U"], dtype=np.float64)
    S = np.array(d["S"], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, f"trajectory{i}.npz"), X=X, U=U, S=S)
    print(i, X.shape, U.shape, S.shape, "s_max", X[-1, 0])
