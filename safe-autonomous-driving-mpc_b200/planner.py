"""Host-side mirror of the reference planner's function-evaluation surface on top of the C ABI.

``PlannerEvaluator`` mirrors the parts of TrajectoryOptimizer (trajectory_planning.py:8-349) that SciPy calls over and
over -- ``cost`` (:128-170), the Hermite-Simpson defect closures (:183-208) and the inequality closures (:249-347) --
as batched GPU evaluations over every collocation interval of every chunk, with analytic Jacobian and
Lagrangian-Hessian blocks.  The outer optimisation loop (:351-390, :419-559) stays on the host by design.

``k_ref_fun`` is the curvature column of the TrajectoryLoader table (linear, extrapolating); ``v_min_fun`` /
``v_max_fun`` may be constants or per-node arrays evaluated by the caller.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import PlannerParams, check


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


class PlannerEvaluator:
    def __init__(self, tracker, N, dt=0.3, simpson_sign=-1, s_total=None, v_min=0.0, v_max=None, **weights):
        """tracker: a BatchedTracker (owns the device table and context); N: intervals per chunk."""
        self._t = tracker
        self._lib = tracker._lib
        p = PlannerParams()
        check(self._lib.mpcb_planner_default_params(C.byref(p)))
        p.dt = dt
        p.simpson_sign = simpson_sign
        p.s_total = float(tracker.X_ref.s_max if s_total is None else s_total)
        p.v_min = v_min
        p.v_max = float(tracker.X_ref.X_ref[:, 4].max() if v_max is None else v_max)
        for k, v in weights.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown parameter {k!r}")
            setattr(p, k, v)
        self.params = p
        self.N = int(N)
        self.dt = dt
        self.u_min = np.array(p.u_min[:])
        self.u_max = np.array(p.u_max[:])
        self.k_min, self.k_max, self.a_max = p.k_min, p.k_max, p.a_max
        self.w_y, self.w_s, self.w_u, self.w_slack = p.w_y, p.w_s, p.w_u, p.w_slack

    # ---- reference helpers (:91-126) ----------------------------------------------------------------
    def unpack(self, z):
        N = self.N
        X = z[: 5 * (N + 1)].reshape(N + 1, 5)
        U = z[5 * (N + 1): 5 * (N + 1) + 2 * N].reshape(N, 2)
        S = z[5 * (N + 1) + 2 * N:]
        return X, U, S

    @staticmethod
    def pack(X, U, S):
        return np.concatenate([np.ravel(X), np.ravel(U), np.ravel(S)])

    # ---- device-tensor entry points -----------------------------------------------------------------
    def eval_defects(self, z, lam=None, want_jac=True, want_hess=False, out=None, stream=None):
        """z [C, 8N+5] (torch cuda f64) -> dict(defect [C,N,5], jac [C,N,5,12], hess [C,N,12,12])."""
        import torch
        h = self._t._need()
        Cn, N = z.shape[0], self.N
        assert z.dtype == torch.float64 and z.is_contiguous() and z.shape[1] == 8 * N + 5
        dev = z.device
        if out is None:
            out = dict(defect=torch.empty((Cn, N, 5), dtype=torch.float64, device=dev))
            if want_jac:
                out["jac"] = torch.empty((Cn, N, 5, 12), dtype=torch.float64, device=dev)
            if want_hess:
                out["hess"] = torch.empty((Cn, N, 12, 12), dtype=torch.float64, device=dev)
        if want_hess:
            assert lam is not None and lam.dtype == torch.float64 and lam.is_contiguous()
        st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        check(self._lib.mpcb_hs_eval(h, C.byref(self.params), Cn, N, z.data_ptr(),
                                     lam.data_ptr() if want_hess else None, out["defect"].data_ptr(),
                                     out["jac"].data_ptr() if "jac" in out else None,
                                     out["hess"].data_ptr() if "hess" in out else None, C.c_void_p(st)),
              "mpcb_hs_eval")
        return out

    def eval_nodes(self, z, s0, vmin_nodes=None, vmax_nodes=None, stream=None):
        """z [C, 8N+5], s0 [C] (torch cuda f64) -> dict(node_rows [C,N+1,6], ctrl_rows [C,N,5], cost_terms [C,N],
        cost [C], cost_grad [C,8N+5])."""
        import torch
        h = self._t._need()
        Cn, N = z.shape[0], self.N
        dev = z.device
        f = dict(dtype=torch.float64, device=dev)
        out = dict(node_rows=torch.empty((Cn, N + 1, 6), **f), ctrl_rows=torch.empty((Cn, N, 5), **f),
                   cost_terms=torch.empty((Cn, N), **f), cost=torch.empty((Cn,), **f),
                   cost_grad=torch.empty((Cn, 8 * N + 5), **f))
        st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        check(self._lib.mpcb_hs_nodes(h, C.byref(self.params), Cn, N, z.data_ptr(), s0.data_ptr(),
                                      vmin_nodes.data_ptr() if vmin_nodes is not None else None,
                                      vmax_nodes.data_ptr() if vmax_nodes is not None else None,
                                      out["node_rows"].data_ptr(), out["ctrl_rows"].data_ptr(),
                                      out["cost_terms"].data_ptr(), out["cost"].data_ptr(),
                                      out["cost_grad"].data_ptr(), C.c_void_p(st)), "mpcb_hs_nodes")
        return out

    # ---- host-array convenience (copies through the library's own allocators; no torch needed) --------
    def evaluate_host(self, z, lam=None, s0=None, want_jac=True, want_hess=False):
        """numpy in, numpy out: everything mpcb_hs_eval and mpcb_hs_nodes produce."""
        lib, h = self._lib, self._t._need()
        z = np.ascontiguousarray(z, dtype=np.float64)
        if z.ndim == 1:
            z = z[None]
        Cn, N = z.shape[0], self.N
        assert z.shape[1] == 8 * N + 5
        s0 = np.ascontiguousarray(z[:, 0] if s0 is None else s0, dtype=np.float64).reshape(Cn)
        want_hess = want_hess and lam is not None
        ins = dict(z=z, s0=s0)
        if want_hess:
            ins["lam"] = np.ascontiguousarray(lam, dtype=np.float64).reshape(Cn, N, 5)
        outs = dict(defect=np.empty((Cn, N, 5)), node_rows=np.empty((Cn, N + 1, 6)), ctrl_rows=np.empty((Cn, N, 5)),
                    cost_terms=np.empty((Cn, N)), cost=np.empty(Cn), cost_grad=np.empty((Cn, 8 * N + 5)))
        if want_jac:
            outs["jac"] = np.empty((Cn, N, 5, 12))
        if want_hess:
            outs["hess"] = np.empty((Cn, N, 12, 12))
        d = {}
        try:
            for k, a in list(ins.items()) + list(outs.items()):
                p = C.c_void_p()
                check(lib.mpcb_device_alloc(h, C.byref(p), a.nbytes), "mpcb_device_alloc")
                d[k] = p
            for k, a in ins.items():
                check(lib.mpcb_memcpy_h2d(h, d[k], _dp(a), a.nbytes))
            check(lib.mpcb_hs_eval(h, C.byref(self.params), Cn, N, d["z"], d.get("lam"), d["defect"], d.get("jac"),
                                   d.get("hess"), None), "mpcb_hs_eval")
            check(lib.mpcb_hs_nodes(h, C.byref(self.params), Cn, N, d["z"], d["s0"], None, None, d["node_rows"],
                                    d["ctrl_rows"], d["cost_terms"], d["cost"], d["cost_grad"], None),
                  "mpcb_hs_nodes")
            for k, a in outs.items():
                check(lib.mpcb_memcpy_d2h(h, _dp(a), d[k], a.nbytes))     # cudaMemcpy synchronises the null stream
        finally:
            for p in d.values():
                lib.mpcb_device_free(h, p)
        return outs

    # ---- the rows the reference adds around the interval closures (host-side, trivial) -----------------
    def boundary_rows(self, z, x0, s_target, is_final_chunk):
        """initial_x0 (:214-218) and terminal rows (:221-246) of one chunk."""
        X, _, _ = self.unpack(np.asarray(z))
        init = X[0] - x0
        if is_final_chunk:
            return init, np.array([X[self.N, 0] - s_target, X[self.N, 4]]), "eq"
        return init, np.array([X[self.N, 0] - s_target / 2]), "ineq"
