"""Callers of the boundary, host side: obstacle-set producer and closed-loop driver.

``ObstaclesFSM``   mirrors trajectory_tracking.py:266-374 (two independent state machines -> list of
                   {'s','v','type'} dicts); scenario constants are constructor arguments instead of edits to the
                   source (the reference keeps the trajectory3 set in a commented block, :313-327).
``run_simulation`` mirrors trajectory_tracking.py:377-443 minus printing and plotting; returns the same tuple.
"""
import numpy as np

SCENARIO_TRAJECTORY2 = dict(obs_trigger_s=710.0, obs_start_s=780.0, obs_v=4.0, obs_end_s=1050.0,
                            tl_pos=550.0, tl_trigger_s=100.0, tl_stop_duration=20.0)
SCENARIO_TRAJECTORY3 = dict(obs_trigger_s=5.0, obs_start_s=150.0, obs_v=4.0, obs_end_s=850.0,
                            tl_pos=2000.0, tl_trigger_s=100.0, tl_stop_duration=20.0)


class ObstaclesFSM:
    def __init__(self, dynamic_obstacle=False, traffic_light=False, scenario=None):
        sc = dict(SCENARIO_TRAJECTORY2)
        if scenario:
            sc.update(scenario)
        self.dynamic_obstacle = dynamic_obstacle
        self.traffic_light = traffic_light
        for k, v in sc.items():
            setattr(self, k, v)
        self.obs_active = False
        self.obs_s = self.obs_start_s
        self.obs_has_triggered = False
        self.tl_state = "RED"
        self.tl_timer = 0.0
        self.tl_waiting = False

    def update(self, dt, s, v):
        active = []
        if self.dynamic_obstacle:
            if s >= self.obs_trigger_s and not self.obs_has_triggered:
                self.obs_has_triggered = self.obs_active = True
            if self.obs_active:
                self.obs_s += self.obs_v * dt
                if self.obs_s > self.obs_end_s:
                    self.obs_active = False
                else:
                    active.append({"s": self.obs_s, "v": self.obs_v, "type": "car"})
        if self.traffic_light and self.tl_state == "RED":
            gap = self.tl_pos - s
            if 0 < gap < self.tl_trigger_s:
                active.append({"s": self.tl_pos, "v": 0.0, "type": "light"})
                if v < 0.1 and gap < 10.0:
                    self.tl_waiting = True
            if self.tl_waiting:
                self.tl_timer += dt
                if self.tl_timer >= self.tl_stop_duration:
                    self.tl_state, self.tl_waiting = "GREEN", False
        return active, self.tl_state


def run_simulation(mpc, fsm, trajectory, max_steps=200000, record_flags=None):
    """Closed loop of trajectory_tracking.py:377-443: FSM -> solve -> explicit-Euler plant step."""
    x = np.array([0.0, 0.0, 0.0, 0.0, 0.5])
    cur_s = x[0]
    hist_x, hist_u, hist_t, hist_preds, hist_obs_s, hist_tl = [x], [], [], [], [], []
    step = 0
    while cur_s <= trajectory.s_max - 1.0 and step < max_steps:
        obstacles, tl_state = fsm.update(mpc.dt, x[0], x[4])
        u_opt, pred_X, sec = mpc.solve(x, obstacles)
        k_ref = trajectory.get_state(cur_s)[3]
        x = x + mpc.dt * mpc.dynamics(x, u_opt, k_ref)
        cur_s = x[0]
        hist_x.append(x)
        hist_u.append(u_opt)
        hist_t.append(sec)
        hist_preds.append(pred_X)
        hist_tl.append(tl_state)
        hist_obs_s.append(next((o["s"] for o in obstacles if o["type"] == "car"), np.nan))
        if record_flags is not None:
            record_flags.append((getattr(mpc, "last_status", None), len(obstacles)))
        step += 1
    return np.array(hist_x), np.array(hist_u), np.array(hist_t), hist_preds, hist_obs_s, hist_tl, trajectory
