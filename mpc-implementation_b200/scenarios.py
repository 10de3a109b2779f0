"""Scenario registry: everything that differs between the reference's six Python scripts.

Each reference script is the same NLP with different module-level constants:
    Python/NMPC_TT.py            T=1   3 obstacles r=30  target (12, 0.01)            700 steps
    Python/T_Trajectory.py       T=.2  3 dummy obstacles  target v=13.5, "T" schedule 1633 steps
    Python/Plus Trajectory.py    T=.2  3 dummy obstacles  target v=20, pulse schedule 1223 steps
    Python/Race Trajectory 1.py  T=.2  3 dummy obstacles  target v=14, race schedule  1595 steps
    Python/Race Track 2.py       T=.2  10 obstacles r=50  target v=12, oval           2000 steps
    Python/10_obstacles.py       T=.2  3 real r=100 + 7 dummies, target v=13          1595 steps
and the gimbal-less model variant (model = 1: 5 states, 3 controls, p[8], rows [z, theta], distance-only cost):
    MATLAB/Dynamic Obstacles/NMPC_TT.m  T=.2  no obstacles  target (15, 0.12)          100 steps
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Sequence, Tuple

import numpy as np

PI = math.pi
NX, NU, NP = 8, 6, 11


def _piecewise(v: float, table: Sequence[Tuple[int, float]]) -> Callable[[int], Tuple[float, float]]:
    """`if mpc_iter >= b: con_t = [v, om]` chains of the scripts: the last breakpoint <= i wins."""
    def sched(i: int) -> Tuple[float, float]:
        om = 0.0
        for b, w in table:
            if i >= b:
                om = w
        return (v, om)
    return sched


_Q = (PI / 2) / 12
# T_Trajectory.py:27-57
_T_TABLE = [(100, _Q), (160, 0.0), (260, -_Q), (320, 0.0), (420, _Q), (480, 0.0), (580, _Q), (640, 0.0),
            (740, -_Q), (800, 0.0), (900, _Q), (960, 0.0), (1060, _Q), (1120, 0.0), (1573, _Q)]
# Plus Trajectory.py:25-69 : one-step pulses of +-(pi/2)*5 rad/s
_P = (PI / 2) * 5
_PLUS_TABLE = []
for _b, _s in [(101, 1), (203, -1), (305, 1), (407, 1), (509, -1), (611, 1), (713, 1), (815, -1), (917, 1), (1019, 1), (1121, -1)]:
    _PLUS_TABLE += [(_b, _s * _P), (_b + 1, 0.0)]
# Race Trajectory 1.py:27-57 (10_obstacles.py:30-60 uses the same breakpoints)
_RACE1_TABLE = [(300, -(PI / 2) / 24), (360, 0.0), (410, (PI / 2) / 24), (470, 0.0), (570, ((11 * PI) / 18) / 12), (630, 0.0),
                (780, ((7 * PI) / 18) / 12), (840, 0.0), (940, -(3 * PI / 18) / 12), (1000, 0.0), (1100, (3 * PI / 18) / 12),
                (1160, 0.0), (1335, (PI / 2) / 12), (1395, 0.0), (1535, (PI / 2) / 12)]
# Race Track 2.py:28-36
_RT2_TABLE = [(500, PI / 100), (1000, 0.0), (1500, PI / 100)]

_DUMMY3 = ((10000.0, 10000.0, 30.0),) * 3


@dataclass
class Scenario:
    name: str
    script: str
    T: float
    N: int
    obstacles: Tuple[Tuple[float, float, float], ...]     # (cx, cy, obs_r)
    x_init: Tuple[float, ...]
    target_init: Tuple[float, float, float]
    schedule: Callable[[int], Tuple[float, float]]         # mpc_iter -> (v_target, omega_target)
    steps: int
    uav_r: float = 5.0
    w1: float = 1.0
    w2: float = 2.0
    vfov: float = 1.0
    hfov: float = 1.0
    model: int = 0        # 0: UAV + gimballed camera (every Python script); 1: gimbal-less tracker (MATLAB/Dynamic Obstacles/NMPC_TT.m)

    @property
    def n_obs(self) -> int:
        return len(self.obstacles)

    @property
    def nu(self) -> int:
        return 3 if self.model else NU

    @property
    def n_box(self) -> int:
        return 2 if self.model else 5

    @property
    def n_p(self) -> int:
        return 8 if self.model else NP

    @property
    def n_w(self) -> int:
        return self.nu * self.N

    @property
    def n_g(self) -> int:
        return (self.n_box + self.n_obs) * (self.N + 1)

    def obstacle_table(self) -> np.ndarray:
        """[n_obs][3] = cx, cy, UAV_r + obs_r  (the constants of NMPC_TT.py:241-243)."""
        o = np.asarray(self.obstacles, dtype=np.float64).reshape(-1, 3).copy()
        o[:, 2] += self.uav_r
        return o

    def bounds(self):
        """lbx, ubx, lbg, ubg exactly as NMPC_TT.py:62-89, :269-306 (Race Track 2.py:289-341); model 1: the first three
        controls and the first two rows of the same tables (MATLAB/Dynamic Obstacles/NMPC_TT.m:14-22, :127-133)."""
        N = self.N
        lo = [14.0, -PI / 30, -PI / 21, -PI / 30, -PI / 30, -PI / 30][:self.nu]
        hi = [30.0, PI / 30, PI / 21, PI / 30, PI / 30, PI / 30][:self.nu]
        lbx = np.tile(np.array(lo), N)
        ubx = np.tile(np.array(hi), N)
        glo = [75.0, -0.2618, -PI / 6, -PI / 6, -PI / 2][:self.n_box] + [-np.inf] * self.n_obs
        ghi = [150.0, 0.2618, PI / 6, PI / 6, PI / 2][:self.n_box] + [0.0] * self.n_obs
        lbg = np.tile(np.array(glo), N + 1)
        ubg = np.tile(np.array(ghi), N + 1)
        return lbx, ubx, lbg, ubg

    def with_horizon(self, N: int) -> "Scenario":
        import dataclasses
        return dataclasses.replace(self, N=N, name=f"{self.name}_N{N}")


_X99 = (99.0, 150.0, 80.0, 0.0, 0.0, 0.0, 0.0, 0.0)
_TGT = (100.0, 150.0, 0.0)

SCENARIOS = {
    "nmpc_tt": Scenario("nmpc_tt", "Python/NMPC_TT.py", 1.0, 15,
                        ((175.0, 820.0, 30.0), (-134.0, 155.0, 30.0), (441.0, 343.0, 30.0)),
                        (90.0, 150.0, 80.0, 0.0, 0.0, 0.0, 0.0, 0.0), _TGT, lambda i: (12.0, 0.01), 700),
    "t_trajectory": Scenario("t_trajectory", "Python/T_Trajectory.py", 0.2, 15, _DUMMY3, _X99, _TGT,
                             _piecewise(13.5, _T_TABLE), 1633),
    "plus_trajectory": Scenario("plus_trajectory", "Python/Plus Trajectory.py", 0.2, 15, _DUMMY3, _X99, _TGT,
                                _piecewise(20.0, _PLUS_TABLE), 1223),
    "race_trajectory_1": Scenario("race_trajectory_1", "Python/Race Trajectory 1.py", 0.2, 15, _DUMMY3, _X99, _TGT,
                                  _piecewise(14.0, _RACE1_TABLE), 1595),
    "race_track_2": Scenario("race_track_2", "Python/Race Track 2.py", 0.2, 15,
                             tuple((float(a), float(b), 50.0) for a, b in
                                   [(0, 80), (500, 245), (1000, 70), (1500, 295), (1765, 550), (1500, 750), (1000, 1005),
                                    (500, 800), (-100, 950), (-200, 550)]),
                             _X99, _TGT, _piecewise(12.0, _RT2_TABLE), 2000),
    "10_obstacles": Scenario("10_obstacles", "Python/10_obstacles.py", 0.2, 15,
                             ((500.0, 20.0, 100.0), (1700.0, 197.0, 100.0), (130.0, 830.0, 100.0)) + ((10000.0, 10000.0, 100.0),) * 7,
                             _X99, _TGT, _piecewise(13.0, _RACE1_TABLE), 1595),
    # model variant: x0 / xs NMPC_TT.m:138-139, constant target input shift1.m:9, sim_time / T = 100 steps :144
    "gimbal_less": Scenario("gimbal_less", "MATLAB/Dynamic Obstacles/NMPC_TT.m", 0.2, 15, (), (90.0, 150.0, 80.0, 0.0, 0.0),
                            _TGT, lambda i: (15.0, 0.12), 100, w1=1.0, w2=0.0, model=1),
}


def get(name: str) -> Scenario:
    return SCENARIOS[name]


# ------------------------------------------------------------------------------------------------
# synthetic Monte-Carlo instances: randomised UAV states / targets that keep
# every stage-0 row of g strictly feasible and break the mirror symmetry of the scripts' first solve
# ------------------------------------------------------------------------------------------------
def random_instances(sc: Scenario, B: int, seed: int):
    """Returns p [B,11] (model 1: [B,8], the same draws without the camera states), target (v, omega) [B,2] as float64 numpy arrays."""
    rng = np.random.default_rng(seed)
    x = np.empty((B, NX))
    cx, cy = sc.x_init[0], sc.x_init[1]
    x[:, 0] = cx + rng.uniform(-50, 50, B)
    x[:, 1] = cy + rng.uniform(-50, 50, B)
    x[:, 2] = rng.uniform(80, 145, B)
    x[:, 3] = rng.uniform(-0.2, 0.2, B)
    x[:, 4] = rng.uniform(-PI, PI, B)
    x[:, 5] = rng.uniform(-0.4, 0.4, B)
    x[:, 6] = rng.uniform(-0.4, 0.4, B)
    x[:, 7] = rng.uniform(-1.0, 1.0, B)
    tgt = np.empty((B, 3))
    tgt[:, 0] = sc.target_init[0] + rng.uniform(-50, 50, B)
    tgt[:, 1] = sc.target_init[1] + rng.uniform(-50, 50, B)
    tgt[:, 2] = rng.uniform(-PI, PI, B)
    vw = np.empty((B, 2))
    vw[:, 0] = rng.uniform(8, 20, B)
    vw[:, 1] = rng.uniform(-0.05, 0.05, B)
    return np.concatenate([x[:, :5] if sc.model else x, tgt], axis=1), vw
