#!/usr/bin/env python
"""State checksums after a fixed episode (dev tool for A/B builds: run once per ORCA_B200_LIB and diff)."""
import hashlib
import sys

import torch

sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), ".."))
from collision_avoidance_b200 import _lib, envs, scenarios  # noqa: E402
from collision_avoidance_b200.sim import BatchedRVOSimulator  # noqa: E402


def h(t):
    return hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


env = envs.Collision_Avoidance_Env(numAgents=10, num_envs=20000, seed=4)
theta = (torch.rand(20000, 10, device="cuda", generator=torch.Generator("cuda").manual_seed(5)) - 0.5) * 0.6
for _ in range(220):
    obs, rew, done, _ = env.step(theta)
print("env10", h(env.sim.pos), h(env.sim.vel), h(obs), h(rew), env.sim.read_stats())
for (E, N, k) in ((4096, 16, 10), (4096, 13, 5), (4096, 7, 3)):
    scn = scenarios.circle(E, N, seed=9)
    scn.params = dict(scn.params, maxNeighbors=k)
    sim = BatchedRVOSimulator(E, N, device="cuda:0", **scn.params)
    sim.set_obstacles(scn.obstacles, per_env=scn.per_env_obstacles)
    sim.pos.copy_(torch.from_numpy(scn.pos))
    sim.vel.copy_(torch.from_numpy(scn.vel))
    goal = torch.from_numpy(scn.goal).cuda()
    for _ in range(300):
        sim.env_step(policy=_lib.POLICY_GOAL, goal=goal)
    print(f"circle{N}k{k}", h(sim.pos), h(sim.vel), sim.read_stats())
