#!/usr/bin/env python
"""Device-timed throughput of every GPU config of BASELINE.json on ONE GPU's share of it
(dev/measurement tool; `bench.py` stays the contract line for configs[1]).

  cfg2  circle 16 agents x 65,536 envs, ORCA policy
  cfg3  circle 32 agents x 32,768 envs (the per-GPU share of 262,144 over 8), ALAN, 8 actions
  cfg4  crowd 256 agents + 4 blocks x 2,048 envs (share of 16,384 over 8), ORCA policy
  cfg5  crowd 1,000,000 agents, one env, uniform grid
  env   default gym world 10 agents x 100,000 envs, RL step + laser observation; and the closed
        loop observation -> policy network -> step
Prints one JSON line per config: agent-steps/s, us/step, algorithmic-byte HBM fraction."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from collision_avoidance_b200 import _lib, alan, envs, scenarios  # noqa: E402
from collision_avoidance_b200.sim import BatchedRVOSimulator  # noqa: E402

PEAK = 6553.9
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def timed(step, steps, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def orca_policy(scn):
    E, N = scn.num_envs, scn.agents_per_env
    sim = BatchedRVOSimulator(E, N, **scn.params)
    sim.set_obstacles(scn.obstacles, per_env=scn.per_env_obstacles)
    sim.pos.copy_(torch.from_numpy(scn.pos))
    sim.vel.copy_(torch.from_numpy(scn.vel))
    st = dict(goal=torch.from_numpy(scn.goal).cuda(), goal2=torch.from_numpy(scn.goal2).cuda(),
              agent_done=torch.zeros(E, N, dtype=torch.uint8, device="cuda"), arrival_time=torch.zeros(E, N, device="cuda"),
              env_step=torch.zeros(E, dtype=torch.int32, device="cuda"),
              env_done_cnt=torch.zeros(E, dtype=torch.int32, device="cuda"))
    stats = not os.environ.get("BC_NO_STATS")
    return sim, lambda: sim.env_step(policy=_lib.POLICY_GOAL, done_mode=_lib.DONE_GOAL_RADIUS_DEFERRED, collect_stats=stats, **st)


def report(name, agents, ms, bytes_per, extra=None):
    v = agents / (ms * 1e-3)
    line = {"config": name, "agent_steps_per_s": v, "us_per_step": ms * 1e3, "bytes_per_agent_step": bytes_per,
            "hbm_frac_algorithmic": bytes_per * v / 1e9 / PEAK}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg4", "cfg5", "env"]
    steps, warmup = int(os.environ.get("BC_STEPS", 200)), int(os.environ.get("BC_WARMUP", 20))
    if "cfg2" in which:
        sim, step = orca_policy(scenarios.circle(65536, 16, seed=1234))
        report("cfg2 circle16 x65536 ORCA", 65536 * 16, timed(step, steps, warmup), 41, sim.read_stats())
        del sim, step
    if "cfg3" in which:
        s = alan.Collision_Avoidance_Sim(numAgents=32, scenario="circle", num_envs=32768, seed=1)
        report("cfg3 circle32 x32768 ALAN(8 actions)", 32768 * 32, timed(s.online_step, steps, warmup), 82,
               s.sim.read_stats())
        del s
    if "cfg4" in which:
        sim, step = orca_policy(scenarios.crowd(2048, 256, seed=2, blocks=4))
        report("cfg4 crowd256+4 blocks x2048 ORCA", 2048 * 256, timed(step, steps, warmup), 48, sim.read_stats())
        del sim, step
    if "cfg5" in which:
        sim, step = orca_policy(scenarios.crowd(1, 1_000_000, seed=3))
        report("cfg5 crowd 1M agents x1 ORCA (grid)", 1_000_000, timed(step, steps, warmup), 150, sim.read_stats())
        del sim, step
    if "env" in which:
        E, N = 100_000, 10
        env = envs.Collision_Avoidance_Env(numAgents=N, num_envs=E, seed=4)
        theta = (torch.rand(E, N, device="cuda", generator=torch.Generator("cuda").manual_seed(5)) - 0.5) * 0.6
        report("gym env 10 agents x100000 RL step + obs", E * N, timed(lambda: env.step(theta), steps, warmup), 45 + 256,
               env.sim.read_stats())
        # the closed loop of the RL shell: observation -> shared policy network (tcgen05) -> fused step
        from collision_avoidance_b200.policy import SharedMLPPolicy
        pol = SharedMLPPolicy(env.sim, seed=1)
        state = {"obs": env._get_obs()}

        def loop_step():
            state["obs"] = env.step(pol.act(state["obs"]))[0]
        report("gym env 10 agents x100000 RL loop: obs + policy net + step", E * N, timed(loop_step, steps, warmup),
               45 + 256 + 264, env.sim.read_stats())


if __name__ == "__main__":
    main()
