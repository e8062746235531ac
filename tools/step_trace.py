#!/usr/bin/env python
"""Per-100-step throughput trace of the fused step over a whole episode (dev tool).
usage: python tools/step_trace.py [scenario] [E] [N] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from collision_avoidance_b200 import _lib, scenarios  # noqa: E402
from collision_avoidance_b200.sim import BatchedRVOSimulator  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "circle"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1200
chunk = int(sys.argv[5]) if len(sys.argv) > 5 else 100
kw = {"blocks": 4} if name == "crowd_blocks" else {}
scn = scenarios.make("crowd" if name == "crowd_blocks" else name, E, N, seed=0, **kw)
sim = BatchedRVOSimulator(E, N, **scn.params)
sim.set_obstacles(scn.obstacles, per_env=scn.per_env_obstacles)
sim.pos.copy_(torch.from_numpy(scn.pos))
sim.vel.copy_(torch.from_numpy(scn.vel))
goal = torch.from_numpy(scn.goal).cuda()
goal2 = torch.from_numpy(scn.goal2).cuda()
done = torch.zeros(E, N, dtype=torch.uint8, device="cuda")
arr = torch.zeros(E, N, device="cuda")
estep = torch.zeros(E, dtype=torch.int32, device="cuda")
dcnt = torch.zeros(E, dtype=torch.int32, device="cuda")
for c in range(steps // chunk):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(chunk):
        sim.env_step(policy=_lib.POLICY_GOAL, goal=goal, goal2=goal2, done_mode=_lib.DONE_GOAL_RADIUS, agent_done=done,
                     arrival_time=arr, env_step=estep, env_done_cnt=dcnt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / chunk
    st = sim.read_stats()
    print(f"steps {c * chunk:5d}-{c * chunk + chunk - 1:5d}: {ms * 1000:8.1f} us/step {E * N / ms * 1e3:.3e} agent-steps/s "
          f"lp3={st['lp3_calls']} coll={st['collisions']} fin={st['finished']} ovf={st['overflow']}", flush=True)
