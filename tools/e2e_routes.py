#!/usr/bin/env python
"""Wall time per orca_step_host call of each route (dev tool): cfg2, pinned host buffers.
usage: tools/e2e_routes.py [route:chunks ...]   default: direct staged:2 staged:4 staged:8"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from collision_avoidance_b200 import _lib, scenarios  # noqa: E402
from collision_avoidance_b200.sim import BatchedRVOSimulator  # noqa: E402

E, N = 65536, 16
scn = scenarios.circle(E, N, seed=1234)
for spec in sys.argv[1:] or ["direct", "staged:2", "staged:4", "staged:8"]:
    route, _, chunks = spec.partition(":")
    os.environ["ORCA_B200_HOST_NO_AUTOTUNE"] = "1"
    if route == "staged":
        os.environ["ORCA_B200_HOST_NO_MAPPED"] = "1"
    else:
        os.environ.pop("ORCA_B200_HOST_NO_MAPPED", None)
    if chunks:
        os.environ["ORCA_B200_HOST_CHUNKS"] = chunks
    else:
        os.environ.pop("ORCA_B200_HOST_CHUNKS", None)
    sim = BatchedRVOSimulator(E, N, **scn.params)
    sim.set_obstacles(scn.obstacles)
    pos_h = torch.from_numpy(scn.pos.copy()).pin_memory()
    vel_h = torch.from_numpy(scn.vel.copy()).pin_memory()
    goal_h = torch.from_numpy(scn.goal.copy()).pin_memory()
    sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=True, steps=20)
    for _ in range(5):
        sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 200
    for _ in range(K):
        sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
    dt = (time.perf_counter() - t0) / K
    print(f"{spec:12s} {dt * 1e3:.3f} ms/call  {E * N / dt:.3e} agent-steps/s  checksum {float(pos_h.sum()):.6f}", flush=True)
    del sim
