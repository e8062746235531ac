import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from collision_avoidance_b200 import _lib, scenarios
from collision_avoidance_b200.sim import BatchedRVOSimulator
E, N = 65536, 16
scn = scenarios.circle(E, N, seed=1234)
sim = BatchedRVOSimulator(E, N, **scn.params)
sim.set_obstacles(scn.obstacles)
pos_h = torch.from_numpy(scn.pos.copy()).pin_memory()
vel_h = torch.from_numpy(scn.vel.copy()).pin_memory()
goal_h = torch.from_numpy(scn.goal.copy()).pin_memory()
sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=True, steps=150)
for _ in range(12):
    sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
for rep in range(8):
    ts = []
    for _ in range(100):
        t0 = time.perf_counter()
        sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"rep {rep}: mean {sum(ts)/len(ts)*1e3:.3f} ms  min {ts[0]*1e3:.3f}  median {ts[50]*1e3:.3f}  p90 {ts[90]*1e3:.3f}  max {ts[-1]*1e3:.3f}", flush=True)
