#!/usr/bin/env python
"""Observation kernel alone on the gym world (dev tool): microseconds per launch."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from collision_avoidance_b200 import envs
E, N = 100_000, 10
env = envs.Collision_Avoidance_Env(numAgents=N, num_envs=E, seed=4)
theta = (torch.rand(E, N, device="cuda", generator=torch.Generator("cuda").manual_seed(5)) - 0.5) * 0.6
for _ in range(120):
    env.step(theta)
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env._get_obs(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print(os.environ.get("ORCA_B200_LIB", "default"), "obs us:", sum(ts) / len(ts))
