"""Batched gym-style environment: ``Collision_Avoidance_Env`` over many worlds at once.

Mirrors collision_avoidance/envs/collision_avoidence_env.py (``reset`` :461, ``step`` :367,
``orca_step`` :447, ``done_test`` :352, ``_get_obs`` :231) with the same names, argument
meaning and return layout.  gym / ray base classes and the Tk canvas are out of scope (SURVEY
section 2, #7-#9); spaces are described by plain attributes.  With ``num_envs == 1`` the
``'agent_i'``-keyed dict views reproduce the reference's return values one to one; with more
worlds everything is a ``[num_envs, numAgents, ...]`` CUDA tensor.

One ``step`` is two kernel launches for all worlds: the fused step (action rotation, doStep,
reward, done test) and the laser-scan observation.
"""
from __future__ import annotations

from math import pi
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, scenarios, spaces
from .sim import BatchedRVOSimulator


class Collision_Avoidance_Env:
    metadata = {"render.modes": []}

    def __init__(self, numAgents: int = 10, num_envs: int = 1, seed: int = 0, device="cuda:0",
                 reference_rng: bool = False):
        """``reference_rng``: world ``e`` draws its spawn positions from ``random.Random(seed + e)``
        in the reference's call order, i.e. it is the world the reference builds (and re-draws at
        every ``reset``) after ``random.seed(seed + e)``."""
        # constants, collision_avoidence_env.py:27-44
        self.timeStep = 1 / 60.
        self.neighborDist = 1.5
        self.maxNeighbors = 5
        self.timeHorizon = 1.5
        self.radius = 0.5
        self.maxSpeed = 1
        self.laser_num = 16
        self.circle_approx_num = 8
        self.numAgents = int(numAgents)
        self.num_envs = int(num_envs)
        self.envsize = 10
        self.step_count = 0
        self.max_step = 1000
        self.reward_scale = 0.3          # :396
        self.done_x = 2.0                # :359
        # spaces (:52-53): Box(-pi, pi, (1,)) and Box(-nd, nd, (laser_num*4,)), per agent
        self.action_space, self.observation_space = spaces.env_spaces(self.neighborDist, self.laser_num)
        self.action_low, self.action_high, self.action_shape = -pi, pi, (1,)
        self.observation_low, self.observation_high = -self.neighborDist, self.neighborDist
        self.observation_shape = (self.laser_num * 4,)
        self.device = torch.device(device)
        self._rng = np.random.default_rng(seed)
        self._seed = seed
        self._streams = scenarios.reference_streams(seed, self.num_envs) if reference_rng else None
        self._init_world()
        self.reset()

    # ------------------------------------------------------------------ world (:77-123)
    def _init_world(self):
        E, N = self.num_envs, self.numAgents
        if self._streams is not None:
            scn = scenarios.default_env(E, N, reference_rng=self._streams)
        else:
            scn = scenarios.default_env(E, N, seed=int(self._rng.integers(0, 2 ** 31 - 1)))
        self.scn = scn
        dev = self.device
        self.sim = BatchedRVOSimulator(E, N, device=dev, **scn.params)
        self.sim.set_obstacles(scn.obstacles)
        self.sim.pos.copy_(torch.from_numpy(scn.pos))
        self.sim.vel.copy_(torch.from_numpy(scn.vel))       # initial velocity = random unit vector (Q2)
        self.targets_pos = torch.from_numpy(scn.goal).to(dev)       # (1, 5)   :94
        self.targets_done = torch.from_numpy(scn.goal2).to(dev)     # (-10, 5) :361
        self.agents_done = torch.zeros(E, N, dtype=torch.uint8, device=dev)
        self.env_step = torch.zeros(E, dtype=torch.int32, device=dev)
        self.env_done_cnt = torch.zeros(E, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(E, N, dtype=torch.float32, device=dev)
        self.obs = torch.zeros(E, N, self.laser_num * 4, dtype=torch.float32, device=dev)
        self._theta = torch.zeros(E, N, dtype=torch.float32, device=dev)
        # neighbor lists start empty (first construction, Q3) -> all-zero observation
        self.sim._alloc_neighbor_outputs()

    # ------------------------------------------------------------------ gym surface
    def reset(self, env_mask: Optional[torch.Tensor] = None):
        """:461-488.  Re-draws the positions only: velocities, targets of agents that already
        finished and the neighbor lists are NOT reset.  ``env_mask`` ([E] bool)
        restricts the reset to some worlds (vector-env use); default all."""
        E, N = self.num_envs, self.numAgents
        if self._streams is not None:
            # only the selected worlds consume their stream, like separate reference envs would
            sel = range(E) if env_mask is None else [e for e in range(E) if bool(env_mask[e])]
            drawn = scenarios.default_env_reset_positions([self._streams[e] for e in sel], N, self.envsize)
            full = self.sim.pos.cpu().numpy().copy()
            full[list(sel)] = drawn
            new_pos = torch.from_numpy(full).to(self.device)
        else:
            x = self._rng.uniform(self.envsize * 0.5, self.envsize, size=(E, N))
            y = self._rng.uniform(0, self.envsize, size=(E, N))
            new_pos = torch.from_numpy(np.stack([x, y], -1).astype(np.float32)).to(self.device)
        if env_mask is None:
            self.sim.pos.copy_(new_pos)
            self.env_step.zero_()
            self.agents_done.zero_()
            self.env_done_cnt.zero_()
        else:
            m = env_mask.to(self.device).bool()
            self.sim.pos[m] = new_pos[m]
            self.env_step[m] = 0
            self.agents_done[m] = 0
            self.env_done_cnt[m] = 0
        self.step_count = 0
        return self._get_obs()

    def _as_theta(self, action) -> torch.Tensor:
        if isinstance(action, dict):  # {'agent_i': theta}, single world (:373)
            if self.num_envs != 1:
                raise ValueError("dict actions address a single world; pass a [E, N] tensor for a batch")
            vals = [float(np.asarray(action["agent_" + str(i)]).reshape(-1)[0]) for i in range(self.numAgents)]
            self._theta.copy_(torch.tensor(vals, dtype=torch.float32).reshape(1, -1))
            return self._theta
        t = torch.as_tensor(action, dtype=torch.float32, device=self.device).reshape(self.num_envs, self.numAgents)
        self._theta.copy_(t)
        return self._theta

    def step(self, action):
        """:367-416.  Returns (obs, reward, done, info): tensors [E,N,64], [E,N], [E] bool, {}.
        For a dict action (single world) the reference's dict layout is returned instead."""
        theta = self._as_theta(action)
        self.sim.env_step(policy=_lib.POLICY_RL, goal=self.targets_pos, goal2=self.targets_done,
                          done_mode=_lib.DONE_X_BELOW, action_theta=theta, rl_reward_scale=self.reward_scale,
                          done_x_threshold=self.done_x, reward=self.reward, agent_done=self.agents_done,
                          env_step=self.env_step, env_done_cnt=self.env_done_cnt, want_neighbors=True)
        self.step_count += 1
        done = self.done_test() | (self.env_step >= self.max_step)
        obs = self._get_obs()
        if isinstance(action, dict):
            return self.obs_dict(), self.reward_dict(), self.done_dict(done), {("agent_" + str(i)): {} for i in
                                                                              range(self.numAgents)}
        return obs, self.reward, done, {}

    def orca_step(self, action=None):
        """:447-450 (ORCA-only stepping used by the reference's __main__): doStep, goal-directed
        preferred velocity, observation."""
        self.sim.env_step(policy=_lib.POLICY_GOAL, goal=self.targets_pos, env_step=self.env_step,
                          want_neighbors=True)
        self.step_count += 1
        return self._get_obs()

    def done_test(self) -> torch.Tensor:
        """:352-365; the per-agent test ran inside the fused step."""
        return self.env_done_cnt >= self.numAgents

    def _get_obs(self) -> torch.Tensor:
        return self.sim.observe(self.targets_pos, self.obs, self.laser_num, self.circle_approx_num)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def render(self, mode="human"):
        raise NotImplementedError("rendering (Tk) is out of scope")

    def close(self):
        self.sim.close()

    # ------------------------------------------------------------------ dict views (single world)
    def obs_dict(self, env: int = 0) -> Dict[str, list]:
        o = self.obs[env].cpu().numpy()
        return {"agent_" + str(i): o[i].astype(float).tolist() for i in range(self.numAgents)}

    def reward_dict(self, env: int = 0) -> Dict[str, float]:
        r = self.reward[env].cpu().numpy()
        return {"agent_" + str(i): float(r[i]) for i in range(self.numAgents)}

    def done_dict(self, done: torch.Tensor, env: int = 0) -> Dict[str, bool]:
        d = {"__all__": bool(done[env].item())}
        for i in range(self.numAgents):
            d["agent_" + str(i)] = False  # the reference never sets per-agent dones (:467)
        return d
