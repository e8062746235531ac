"""Batched ALAN simulator shell: ``Collision_Avoidance_Sim`` over many worlds at once.

Mirrors collision_avoidance/ALAN/ALAN_true.py (``Collision_Avoidance_Sim``: ``reset`` :79,
``run_sim`` :106, ``online_step`` :569, ``orca_step`` :631, ``done_test`` :547, TTime :125-131,
min TTime :161-172) with the same names and argument meaning.  Differences, all forced by
batching: every per-world scalar becomes a ``[num_envs]`` tensor, the Tk visualisation is
gone, and the unseeded global RNGs are replaced by a seed.

One ``online_step`` / ``orca_step`` is ONE kernel launch for all worlds: action selection,
preferred velocity, doStep, reward, bandit update and the done test are fused
(csrc/orca_step_small.cuh).
"""
from __future__ import annotations

from math import atan2, cos, sin
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, scenarios
from .sim import BatchedRVOSimulator

DEFAULT_ONLINE_ACTIONS = [(1, 0), (0.70711, 0.70711), (0, 1), (-0.70711, 0.70711), (-1, 0), (-0.70711, -0.70711),
                          (0, -1), (0.70711, -0.70711)]  # ALAN_true.py:31-38


def alan_window_steps(time_step: float, timewindow: float) -> int:
    """Steps between the global resets of the action weights.  The reference advances a float64
    timer by ``timeStep`` per step and resets once it is ``>= timewindow`` (ALAN_true.py:621-625);
    all timers move in lock step, so the reset has a fixed period: 121 steps for 1/60 s and 2 s
    (SURVEY Q7).  Computed by replaying that float64 accumulation."""
    t, n = 0.0, 0
    while True:
        t += time_step
        n += 1
        if t >= timewindow:
            return n
        if n > 10_000_000:
            raise ValueError("timewindow is never reached")


def unit_actions(actions: Sequence) -> np.ndarray:
    """(cos, sin) of atan2(action): only the direction of an action is used (ALAN_true.py:592-595)."""
    return np.asarray([(cos(atan2(a[1], a[0])), sin(atan2(a[1], a[0]))) for a in actions], dtype=np.float32)


class Collision_Avoidance_Sim:
    """Batch of ``num_envs`` ALAN worlds of ``numAgents`` agents each."""

    def __init__(self, numAgents: int = 50, scenario: str = "crowd", online_actions: Optional[Sequence] = None,
                 visualize: bool = False, num_envs: int = 1, seed: int = 0, device="cuda:0",
                 reference_rng: bool = False):
        if visualize:
            raise NotImplementedError("the Tk visualisation of the reference is out of scope (SURVEY section 2, #9)")
        # ORCA config, ALAN_true.py:15-20
        self.timeStep = 1 / 60.
        self.neighborDist = 5
        self.maxNeighbors = 10
        self.timeHorizon = 1.5
        self.radius = 0.5
        self.maxSpeed = 1
        # ALAN config, ALAN_true.py:31-49
        self.default_online_actions = list(DEFAULT_ONLINE_ACTIONS)
        self.online_actions = list(online_actions) if online_actions is not None else self.default_online_actions
        self.gamma = 0.6
        self.timewindow = 2
        self.online_temp = 0.2
        # world config, ALAN_true.py:52-62
        self.numAgents = int(numAgents)
        self.scenario = scenario
        self.num_envs = int(num_envs)
        self.seed = int(seed)
        self.reference_rng = bool(reference_rng)   # world e == the reference after random.seed(seed + e)
        self.device = torch.device(device)
        self.max_step = int((10 / self.timeStep) * self.numAgents)
        self.visualize = False
        self._episode = 0
        self.sim: Optional[BatchedRVOSimulator] = None
        self._init_world()

    # ------------------------------------------------------------------ world
    def _init_world(self):
        kw = dict(reference_rng=True) if self.reference_rng else {}
        if self.reference_rng and self.scenario == "circle":
            kw["rotate"] = False
        scn = scenarios.make(self.scenario, self.num_envs, self.numAgents, seed=self.seed + 7919 * self._episode, **kw)
        self.scn = scn
        self.envsize = scn.envsize
        dev = self.device
        if self.sim is None:
            self.sim = BatchedRVOSimulator(self.num_envs, self.numAgents, device=dev, **scn.params)
        self.sim.set_obstacles(scn.obstacles, per_env=scn.per_env_obstacles)
        self.sim.pos.copy_(torch.from_numpy(scn.pos))
        self.sim.vel.copy_(torch.from_numpy(scn.vel))
        self.sim.stats.zero_()
        E, N = self.num_envs, self.numAgents
        self.goal = torch.from_numpy(scn.goal).to(dev)
        self.goal2 = torch.from_numpy(scn.goal2).to(dev)
        self._set_actions(self.online_actions)
        self.agents_done = torch.zeros(E, N, dtype=torch.uint8, device=dev)
        self.agents_time = torch.full((E, N), self.max_step * self.timeStep, dtype=torch.float32, device=dev)
        self.env_step = torch.zeros(E, dtype=torch.int32, device=dev)
        self.env_done_cnt = torch.zeros(E, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(E, N, dtype=torch.float32, device=dev)
        self.action_ids = torch.zeros(E, N, dtype=torch.uint8, device=dev)
        self.step_count = 0
        # min TTime (ALAN_true.py:161-172): straight-line time at max speed, mean + 3 sigma
        d = torch.linalg.norm((self.goal - self.sim.pos).double(), dim=-1) * self.maxSpeed
        self.min_TTime = d.mean(1) + 3 * d.std(1, unbiased=False)
        self.TTime = torch.zeros(E, dtype=torch.float64, device=dev)

    def _set_actions(self, actions):
        """``actions``: one action set for every world (the reference's form), or a list of
        ``num_envs`` action sets -- one per world -- for batched action-space search."""
        per_env = len(actions) > 0 and isinstance(actions[0], (list, tuple)) and len(actions[0]) > 0 and \
            isinstance(actions[0][0], (list, tuple))
        self.online_actions = [list(a) for a in actions] if per_env else list(actions)
        sets = self.online_actions if per_env else [self.online_actions]
        if per_env and len(sets) != self.num_envs:
            raise ValueError("per-world action sets need one set per world")
        A = max(len(a) for a in sets)
        if min(len(a) for a in sets) < 1 or A > _lib.MAX_ACTIONS:
            raise ValueError(f"between 1 and {_lib.MAX_ACTIONS} online actions are supported")
        if per_env:
            table = np.zeros((self.num_envs, A, 2), np.float32)
            table[..., 0] = 1.0
            for e, a in enumerate(sets):
                table[e, :len(a)] = unit_actions(a)
            self.action_table = torch.from_numpy(table).to(self.device)
            self.action_counts = torch.tensor([len(a) for a in sets], dtype=torch.int32, device=self.device)
        else:
            self.action_table = torch.from_numpy(unit_actions(self.online_actions)).to(self.device)
            self.action_counts = None
        self.action_weights = torch.zeros(self.num_envs, self.numAgents, A, dtype=torch.float32, device=self.device)
        self.window_steps = alan_window_steps(self.timeStep, self.timewindow)

    def reset(self, online_actions: Optional[Sequence] = None):
        """ALAN_true.py:79-103: fresh simulator state + (optionally) a new action set."""
        self.online_actions = list(online_actions) if online_actions is not None else self.default_online_actions
        self._episode += 1
        self._init_world()

    # ------------------------------------------------------------------ steps
    def online_step(self, uniforms: Optional[torch.Tensor] = None, steps: int = 1):
        """ALAN_true.py:569-628 for every world.  ``uniforms`` ([E, N] in [0, 1)) overrides the
        in-kernel Philox draw (parity tests).  ``steps`` > 1 issues that many steps back to back
        from C without returning to Python."""
        self.sim.env_step(policy=_lib.POLICY_ALAN, goal=self.goal, goal2=self.goal2, done_mode=_lib.DONE_GOAL_RADIUS,
                          alan_weights=self.action_weights, alan_actions=self.action_table,
                          alan_action_out=self.action_ids, alan_uniform=uniforms,
                          alan_num_actions_env=self.action_counts,
                          alan_window_steps=self.window_steps, alan_gamma=self.gamma, alan_temp=self.online_temp,
                          rng_seed=self.seed * 1_000_003 + self._episode, reward=self.reward,
                          agent_done=self.agents_done, arrival_time=self.agents_time, env_step=self.env_step,
                          env_done_cnt=self.env_done_cnt, steps=steps)

    def orca_step(self, steps: int = 1):
        """ALAN_true.py:631-636 for every world (doStep, then goal-directed preferred velocity).
        The reference updates the preferred velocity BEFORE done_test swaps an arrived agent's
        target (run_sim :116-120), so the step after an arrival still aims at the old goal:
        DONE_GOAL_RADIUS_DEFERRED reproduces that (``agents_done`` holds 2 for that one step)."""
        self.sim.env_step(policy=_lib.POLICY_GOAL, goal=self.goal, goal2=self.goal2,
                          done_mode=_lib.DONE_GOAL_RADIUS_DEFERRED,
                          agent_done=self.agents_done, arrival_time=self.agents_time, env_step=self.env_step,
                          env_done_cnt=self.env_done_cnt, steps=steps)

    def done_test(self) -> torch.Tensor:
        """ALAN_true.py:547-566.  The test itself ran inside the last step; this returns the
        per-world 'all agents done' flags ([E] bool tensor)."""
        return self.env_done_cnt >= self.numAgents

    def run_sim(self, mode: int = 1, max_steps: Optional[int] = None, check_every: int = 64):
        """ALAN_true.py:106-131.  Returns per-world tensors (success, total_time, TTime, min_TTime).
        Worlds that finish early keep being stepped (their arrival times no longer change);
        the host polls for 'all worlds done' every ``check_every`` steps only."""
        if mode not in (0, 1):
            mode = 1
        limit = self.max_step if max_steps is None else int(max_steps)
        done_steps = 0
        while done_steps < limit:
            chunk = min(check_every, limit - done_steps)
            if mode == 1:
                self.online_step(steps=chunk)
            else:
                self.orca_step(steps=chunk)
            done_steps += chunk
            self.step_count += chunk
            if bool(self.done_test().all()):
                break
        self.sim.check_overflow()   # loud, never silent: a dropped obstacle constraint voids the episode
        success = self.done_test()
        times = self.agents_time.double()
        self.TTime = times.mean(1) + 3 * times.std(1, unbiased=False)
        finished_at = times.max(1).values  # step_count * timeStep at the moment the world completed
        total_time = torch.where(success, finished_at,
                                 torch.full_like(finished_at, self.step_count * self.timeStep))
        return success, total_time, self.TTime, self.min_TTime

    # ------------------------------------------------------------------ helpers the reference exposes
    def update_pref_vel(self):
        """ALAN_true.py:483-486: goal-directed preferred velocity into the simulator state."""
        d = self.goal - self.sim.pos
        ang = torch.atan2(d[..., 1].double(), d[..., 0].double())
        self.sim.pref.copy_(torch.stack([torch.cos(ang), torch.sin(ang)], -1).float())

    def comp_pref_vel(self) -> torch.Tensor:
        d = self.goal - self.sim.pos
        ang = torch.atan2(d[..., 1].double(), d[..., 0].double())
        return torch.stack([torch.cos(ang), torch.sin(ang)], -1)
