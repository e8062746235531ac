#!/usr/bin/env python
"""Small batches of every kernel of the library, for `compute-sanitizer`:

  compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
  compute-sanitizer --tool synccheck python tools/sanitize_smoke.py

Shapes follow BASELINE configs 2-5 (tile kernel with ranked selection for 16 / 32 agents, in-block
grid + candidate buffer for 256 agents with per-env obstacle blocks, uniform-grid pipeline) plus
the RL env step with the laser observation and the policy network.  The step kernel re-uses its
shared memory three ways (LP3 queue area <-> in-block grid, dead line columns <-> LP3 programme,
line slots <-> candidate buffer), which is what racecheck is pointed at."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from collision_avoidance_b200 import _lib, alan, envs, scenarios  # noqa: E402
from collision_avoidance_b200.sim import BatchedRVOSimulator  # noqa: E402

STEPS = int(os.environ.get("SAN_STEPS", 4))


def orca(scn, steps=STEPS):
    E, N = scn.num_envs, scn.agents_per_env
    sim = BatchedRVOSimulator(E, N, **scn.params)
    sim.set_obstacles(scn.obstacles, per_env=scn.per_env_obstacles)
    sim.pos.copy_(torch.from_numpy(scn.pos))
    sim.vel.copy_(torch.from_numpy(scn.vel))
    st = dict(goal=torch.from_numpy(scn.goal).cuda(), goal2=torch.from_numpy(scn.goal2).cuda(),
              agent_done=torch.zeros(E, N, dtype=torch.uint8, device="cuda"), arrival_time=torch.zeros(E, N, device="cuda"),
              env_step=torch.zeros(E, dtype=torch.int32, device="cuda"),
              env_done_cnt=torch.zeros(E, dtype=torch.int32, device="cuda"))
    for _ in range(steps):
        sim.env_step(policy=_lib.POLICY_GOAL, done_mode=_lib.DONE_GOAL_RADIUS_DEFERRED, want_neighbors=True, **st)
    torch.cuda.synchronize()
    return sim.read_stats()


def main():
    print("cfg2", orca(scenarios.circle(70, 16, seed=1)))                 # 70 envs: a partly filled last block
    # dense start so that LP3 (block queue + borrowed columns) runs in most blocks
    print("cfg2 dense", orca(scenarios.crowd(40, 16, seed=2)))
    s = alan.Collision_Avoidance_Sim(numAgents=32, scenario="circle", num_envs=20, seed=1)
    s.online_step(steps=STEPS)
    torch.cuda.synchronize()
    print("cfg3", s.sim.read_stats())
    print("cfg4", orca(scenarios.crowd(6, 256, seed=2, blocks=4)))
    print("cfg4 k<K", orca(scenarios.crowd(5, 100, seed=3, blocks=4)))
    print("grid", orca(scenarios.crowd(1, 3000, seed=3), steps=2))
    env = envs.Collision_Avoidance_Env(numAgents=10, num_envs=50, seed=4)
    theta = torch.zeros(50, 10, device="cuda")
    for _ in range(STEPS):
        obs = env.step(theta)[0]
    torch.cuda.synchronize()
    print("env", env.sim.read_stats(), float(obs.abs().sum()))
    from collision_avoidance_b200.policy import SharedMLPPolicy
    pol = SharedMLPPolicy(env.sim, seed=1)
    act = pol.act(obs)
    torch.cuda.synchronize()
    print("policy", tuple(act.shape))
    pos_h = torch.from_numpy(scenarios.circle(64, 16, seed=5).pos.copy()).pin_memory()
    scn = scenarios.circle(64, 16, seed=5)
    sim = BatchedRVOSimulator(64, 16, **scn.params)
    sim.set_obstacles(scn.obstacles)
    vel_h = torch.from_numpy(scn.vel.copy()).pin_memory()
    goal_h = torch.from_numpy(scn.goal.copy()).pin_memory()
    sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=True, steps=2)
    for _ in range(8):
        sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
    print("host", float(pos_h.sum()))


if __name__ == "__main__":
    main()
