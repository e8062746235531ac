"""GPU parity (through the C ABI): one doStep of the CUDA path vs. the CPU oracle on
identical states.  Bar (BASELINE.json north_star): neighbor index sets bit-exact (ties at
equal distance excepted), new velocities / positions within 1e-4 absolute.  The kernels are
built without FMA contraction, so we additionally report (and for these cases require) exact
equality."""
import numpy as np
import pytest

from _common import goal_pref, neighbor_sets_equal_up_to_ties, oracle_sims

pytestmark = pytest.mark.gpu

TOL = 1e-4  # absolute, BASELINE.json north_star


def _gpu_sim(scn, device="cuda:0"):
    import torch
    from collision_avoidance_b200.sim import BatchedRVOSimulator
    sim = BatchedRVOSimulator(scn.num_envs, scn.agents_per_env, device=device, **scn.params)
    sim.set_obstacles(scn.obstacles, per_env=scn.per_env_obstacles)
    sim.pos.copy_(torch.from_numpy(scn.pos))
    sim.vel.copy_(torch.from_numpy(scn.vel))
    return sim


def _run_case(scn, steps):
    """Returns (worst abs diff, fraction of agent-steps that are bit-identical among those whose
    ORDERED neighbor lists agree).  Agents whose lists differ only by the order of bit-equal
    distances (the tie exemption; RVO2's order there depends on its kd-tree permutation) are
    held to the 1e-4 tolerance only."""
    import torch
    sims = oracle_sims(scn)
    gpu = _gpu_sim(scn)
    E, N = scn.num_envs, scn.agents_per_env
    goal = scn.goal
    worst = 0.0
    n_exact = 0
    n_total = 0
    n_ties = 0
    for t in range(steps):
        # shared state: the oracle's current state is loaded into the GPU sim every step
        pos = np.stack([s.positions() for s in sims])
        vel = np.stack([s.velocities() for s in sims])
        pref = goal_pref(pos, goal).astype(np.float32)
        for e, s in enumerate(sims):
            s.set_pref_velocities(pref[e])
            s.doStep()
        gpu.pos.copy_(torch.from_numpy(pos))
        gpu.vel.copy_(torch.from_numpy(vel))
        gpu.pref.copy_(torch.from_numpy(pref))
        nbr_idx, nbr_cnt, onbr_idx, onbr_cnt = [x.cpu().numpy() for x in gpu.neighbors()]
        gpu.doStep()
        gp = gpu.pos.cpu().numpy()
        gv = gpu.vel.cpu().numpy()
        op = np.stack([s.positions() for s in sims])
        ov = np.stack([s.velocities() for s in sims])
        worst = max(worst, float(np.abs(gp - op).max()), float(np.abs(gv - ov).max()))
        for e in range(E):
            for i in range(N):
                o_ids = [x[0] for x in sims[e].agent_neighbors(i)]
                g_ids = list(nbr_idx[e, i, :nbr_cnt[e, i]])
                dsq = lambda j: float(np.float32(((pos[e, i] - pos[e, j]) ** 2).sum()))
                assert neighbor_sets_equal_up_to_ties(o_ids, g_ids, dsq), (t, e, i, o_ids, g_ids)
                o_ob = [x[0] for x in sims[e].obstacle_neighbors(i)]
                g_ob = list(onbr_idx[e, i, :onbr_cnt[e, i]])
                assert o_ob == g_ob, (t, e, i, o_ob, g_ob)
                if o_ids == g_ids:
                    n_total += 1
                    n_exact += int((gv[e, i] == ov[e, i]).all() and (gp[e, i] == op[e, i]).all())
                else:
                    n_ties += 1
    assert worst <= TOL, worst
    print(f"{scn.name}: worst={worst:.3g} exact={n_exact}/{n_total} tie-ordered={n_ties}")
    return worst, n_exact / max(1, n_total)


def test_circle16_matches_oracle():
    from collision_avoidance_b200 import scenarios
    scn = scenarios.circle(8, 16, seed=1)
    worst, exact = _run_case(scn, steps=300)
    assert exact == 1.0, (worst, exact)


def test_circle32_matches_oracle():
    from collision_avoidance_b200 import scenarios
    scn = scenarios.circle(4, 32, seed=2)
    worst, exact = _run_case(scn, steps=300)
    assert exact == 1.0, (worst, exact)


def test_crowd_blocks_matches_oracle():
    from collision_avoidance_b200 import scenarios
    scn = scenarios.crowd(3, 64, seed=3, blocks=4)
    worst, exact = _run_case(scn, steps=200)
    assert exact == 1.0, (worst, exact)


def test_default_env_matches_oracle():
    from collision_avoidance_b200 import scenarios
    scn = scenarios.default_env(6, 10, seed=4)
    worst, exact = _run_case(scn, steps=300)
    assert exact == 1.0, (worst, exact)


def test_deadlock_congested_match_oracle():
    from collision_avoidance_b200 import scenarios
    for scn in (scenarios.deadlock(2, 20, seed=5), scenarios.congested(2, 30, seed=6)):
        worst, exact = _run_case(scn, steps=200)
        assert exact == 1.0, (scn.name, worst, exact)


def test_grid_path_matches_oracle_kdtree():
    """agents_per_env > 256 goes through the uniform-grid pipeline; RVO2 (the oracle) uses its
    kd-tree.  Same neighbor sets, bit-identical velocities."""
    from collision_avoidance_b200 import scenarios
    scn = scenarios.crowd(2, 700, seed=12)
    worst, exact = _run_case(scn, steps=25)
    assert exact == 1.0, (worst, exact)


def test_grid_and_tile_paths_agree_bitwise():
    """The same 256-agent worlds stepped as one 256-agent env (shared-memory tile path) and as
    part of a 512-agent env pair far apart (grid path) give identical bits."""
    import torch
    from collision_avoidance_b200 import scenarios
    from collision_avoidance_b200.sim import BatchedRVOSimulator
    scn = scenarios.crowd(2, 256, seed=13)
    tile = _gpu_sim(scn)
    # one 512-agent world: env 1 shifted by 1000 units so the two crowds never interact
    P = scn.params
    big = BatchedRVOSimulator(1, 512, device="cuda:0", **P)
    shift = np.array([1000.0, 0.0], np.float32)
    pos = np.concatenate([scn.pos[0], scn.pos[1] + shift])[None]
    goal = np.concatenate([scn.goal[0], scn.goal[1] + shift])[None]
    big.pos.copy_(torch.from_numpy(pos))
    big.vel.copy_(torch.from_numpy(np.concatenate([scn.vel[0], scn.vel[1]])[None]))
    wall = scn.obstacles[0]
    big.set_obstacles([wall, [(x + 1000.0, y) for x, y in wall]])
    g_t = torch.from_numpy(scn.goal).cuda()
    g_b = torch.from_numpy(goal).cuda()
    for _ in range(30):
        tile.env_step(policy=1, goal=g_t)
        big.env_step(policy=1, goal=g_b)
    a = tile.vel.cpu().numpy()
    b = big.vel.cpu().numpy()[0]
    # env 0 sits at the same coordinates in both layouts -> identical bits after 30 steps.
    # (env 1 is translated by 1000 units, so its float32 coordinates differ and it is not compared.)
    assert np.array_equal(a[0], b[:256])
    assert np.array_equal(tile.pos.cpu().numpy()[0], big.pos.cpu().numpy()[0][:256])


def test_million_agent_world_neighbor_sets():
    """BASELINE config 5 at full size: 1,000,000 agents, k = 10, uniform grid.  Size-independent
    properties: a random sample of agents has exactly the brute-force k nearest neighbors
    (ascending distance, strict range test), speeds stay <= maxSpeed, nobody is lost."""
    import torch
    from collision_avoidance_b200 import scenarios
    from collision_avoidance_b200.sim import BatchedRVOSimulator
    N = 1_000_000
    scn = scenarios.crowd(1, N, seed=14)
    sim = BatchedRVOSimulator(1, N, device="cuda:0", **scn.params)
    sim.set_obstacles(scn.obstacles)
    sim.pos.copy_(torch.from_numpy(scn.pos))
    sim.vel.copy_(torch.from_numpy(scn.vel))
    goal = torch.from_numpy(scn.goal).cuda()
    for _ in range(3):
        sim.env_step(policy=1, goal=goal)
    pos = sim.pos.clone()
    idx, cnt, _, _, dsq = sim.neighbors(with_distsq=True)
    idx, cnt, dsq = idx.cpu().numpy()[0], cnt.cpu().numpy()[0], dsq.cpu().numpy()[0]
    p = pos.cpu().numpy()[0]
    rng = np.random.default_rng(0)
    nd_sq = np.float32(scn.params["neighborDist"]) ** 2
    for i in rng.integers(0, N, 200):
        d = ((p - p[i]).astype(np.float32) ** 2)
        d2 = (d[:, 0] + d[:, 1]).astype(np.float32)
        d2[i] = np.inf
        cand = np.where(d2 < nd_sq)[0]
        order = cand[np.lexsort((cand, d2[cand]))][:10]
        assert list(order) == list(idx[i, :cnt[i]]), i
        assert np.array_equal(d2[order], dsq[i, :cnt[i]])
    sim.env_step(policy=1, goal=goal)
    v = sim.vel.cpu().numpy()[0]
    assert np.isfinite(v).all()
    # A uniformly random crowd starts with ~8e5 overlapping pairs.  Their "collision" half-planes
    # sit up to 30 speed units from the origin, and RVO2's float32 LP3 intersects nearly parallel
    # ones (|det| just above 1e-5), so a few agents leave the speed disc -- the CPU oracle does
    # exactly the same on such states (bit-identical, see the oracle parity tests).  Everybody
    # else obeys the speed limit.
    speed = np.linalg.norm(v, axis=1)
    assert (speed > 1.0 + 1e-3).mean() < 0.02
    assert sim.read_stats()["overflow"] == 0


def test_grid_equals_tile_over_long_runs(monkeypatch):
    """The same batch stepped by the shared-memory tile path and by the uniform-grid pipeline
    (forced with ORCA_B200_GRID_MIN_AGENTS) stays bit-identical for 150 steps -- including pairs
    that sit within rounding of the neighbor range (cell-size margin) and per-env obstacles."""
    import torch
    from collision_avoidance_b200 import scenarios
    scn = scenarios.crowd(96, 128, seed=21, blocks=4)
    tile = _gpu_sim(scn)
    monkeypatch.setenv("ORCA_B200_GRID_MIN_AGENTS", "2")
    grid = _gpu_sim(scn)
    monkeypatch.delenv("ORCA_B200_GRID_MIN_AGENTS")
    goal = torch.from_numpy(scn.goal).cuda()
    for t in range(150):
        tile.env_step(policy=1, goal=goal)
        grid.env_step(policy=1, goal=goal)
    assert grid.launch_count() == 150 * 5 and tile.launch_count() == 150
    assert torch.equal(tile.pos, grid.pos) and torch.equal(tile.vel, grid.vel)
    assert tile.read_stats() == grid.read_stats()


def test_exact_ties_keep_first_visited_order_on_gpu():
    """Perfectly symmetric rings produce bit-equal distances.  With N <= 10 RVO2's kd-tree is one
    leaf visited in id order, so the ORDERED neighbor lists (not just the sets) must be identical
    to the oracle's, and so must the velocities."""
    import torch
    from collision_avoidance_b200 import scenarios
    for N, k in ((10, 5), (10, 9), (8, 4)):
        scn = scenarios.circle(3, N, seed=3, rotate=False)
        scn.params = dict(scn.params, maxNeighbors=k)
        c = scn.envsize / 2
        d = c - scn.pos
        scn.vel = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
        sims = oracle_sims(scn)
        gpu = _gpu_sim(scn)
        ties = 0
        for _ in range(40):
            pos = np.stack([s.positions() for s in sims])
            vel = np.stack([s.velocities() for s in sims])
            pref = goal_pref(pos, scn.goal).astype(np.float32)
            for e, s in enumerate(sims):
                s.set_pref_velocities(pref[e])
                s.doStep()
            gpu.pos.copy_(torch.from_numpy(pos))
            gpu.vel.copy_(torch.from_numpy(vel))
            gpu.pref.copy_(torch.from_numpy(pref))
            idx, cnt, _, _ = [x.cpu().numpy() for x in gpu.neighbors()]
            gpu.doStep()
            for e in range(3):
                for i in range(N):
                    o = sims[e].agent_neighbors(i)
                    ties += len(set(x[1] for x in o)) < len(o)
                    assert [x[0] for x in o] == list(idx[e, i, :cnt[e, i]]), (N, k, e, i)
            assert np.array_equal(gpu.vel.cpu().numpy(), np.stack([s.velocities() for s in sims]))
        assert ties > 50


@pytest.mark.parametrize("name,small_envs,copies,agents,steps", [("cfg2", 64, 1024, 16, 60), ("cfg4", 16, 128, 256, 25)])
def test_full_size_batches_are_env_independent(name, small_envs, copies, agents, steps):
    """BASELINE configs[1] (65,536 envs x 16 agents) and the per-GPU share of configs[3] (2,048 envs
    x 256 agents with per-env obstacle blocks) at FULL size, through a size-independent property:
    envs never interact, so a batch made of `copies` copies of a few distinct envs must give every
    copy the bits of the small batch -- which the cases above pin to the oracle step by step --
    wherever the copy sits in the batch (block boundaries, last block, 32-bit offsets)."""
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    from collision_avoidance_b200.sim import BatchedRVOSimulator
    if name == "cfg2":
        scn = scenarios.circle(small_envs, agents, seed=21)
    else:
        scn = scenarios.crowd(small_envs, agents, seed=22, blocks=4)
    small = _gpu_sim(scn)
    E = small_envs * copies
    big = BatchedRVOSimulator(E, agents, device="cuda:0", **scn.params)
    if scn.per_env_obstacles:
        big.set_obstacles(list(scn.obstacles) * copies, per_env=True)
    else:
        big.set_obstacles(scn.obstacles, per_env=False)
    big.pos.copy_(torch.from_numpy(np.tile(scn.pos, (copies, 1, 1))))
    big.vel.copy_(torch.from_numpy(np.tile(scn.vel, (copies, 1, 1))))
    g_s = torch.from_numpy(scn.goal).cuda()
    g_b = torch.from_numpy(np.tile(scn.goal, (copies, 1, 1))).cuda()
    # one step from the shared initial state against the oracle (1e-4, north star) ...
    sims = oracle_sims(scn)
    pref = goal_pref(scn.pos, scn.goal).astype(np.float32)
    for e, s in enumerate(sims):
        s.set_pref_velocities(pref[e])
        s.doStep()
    small.env_step(policy=_lib.POLICY_GOAL, goal=g_s)
    big.env_step(policy=_lib.POLICY_GOAL, goal=g_b)
    ov = np.stack([s.velocities() for s in sims])
    assert np.abs(small.vel.cpu().numpy() - ov).max() <= TOL
    # ... then a closed-loop run: every copy keeps the bits of the small batch
    for _ in range(steps):
        small.env_step(policy=_lib.POLICY_GOAL, goal=g_s)
        big.env_step(policy=_lib.POLICY_GOAL, goal=g_b)
    sp, sv = small.pos.cpu().numpy(), small.vel.cpu().numpy()
    bp = big.pos.cpu().numpy().reshape(copies, small_envs, agents, 2)
    bv = big.vel.cpu().numpy().reshape(copies, small_envs, agents, 2)
    assert np.array_equal(bp, np.broadcast_to(sp, bp.shape))
    assert np.array_equal(bv, np.broadcast_to(sv, bv.shape))
    st_s, st_b = small.read_stats(), big.read_stats()
    assert st_b["collisions"] == copies * st_s["collisions"] and st_b["lp3_calls"] == copies * st_s["lp3_calls"]


def test_cfg4_world_256_agents_with_blocks_matches_oracle_for_50_steps():
    """BASELINE configs[3]'s world shape at full agent count: 256 agents + wall + 4 blocks,
    50 oracle-synchronised steps (neighbor lists, obstacle lists, bit-exact velocities)."""
    from collision_avoidance_b200 import scenarios
    scn = scenarios.crowd(2, 256, seed=41, blocks=4)
    worst, exact = _run_case(scn, steps=50)
    assert exact == 1.0, (worst, exact)


def test_million_agent_world_one_oracle_step():
    """BASELINE configs[4] at full size against the oracle itself: one doStep of the 1,000,000
    agent world (uniform grid on the GPU, kd-tree in the oracle), velocities and positions within
    1e-4 and -- wherever the ordered neighbor lists agree -- bit-identical."""
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    from collision_avoidance_b200.sim import BatchedRVOSimulator
    from oracle import rvo2_oracle
    N = 1_000_000
    scn = scenarios.crowd(1, N, seed=15)
    P = scn.params
    sim = BatchedRVOSimulator(1, N, device="cuda:0", **P)
    sim.set_obstacles(scn.obstacles)
    goal = torch.from_numpy(scn.goal).cuda()
    sim.pos.copy_(torch.from_numpy(scn.pos))
    sim.vel.copy_(torch.from_numpy(scn.vel))
    for _ in range(40):            # spread the uniformly random start (heavy overlaps) into a settled crowd
        sim.env_step(policy=_lib.POLICY_GOAL, goal=goal)
    pos, vel = sim.pos.cpu().numpy()[0], sim.vel.cpu().numpy()[0]
    pref = goal_pref(pos, scn.goal[0]).astype(np.float32)
    o = rvo2_oracle.PyRVOSimulator(P["timeStep"], P["neighborDist"], P["maxNeighbors"], P["timeHorizon"],
                                   P["timeHorizonObst"], P["radius"], P["maxSpeed"])
    o.add_agents(pos, vel)
    for poly in scn.obstacles:
        o.addObstacle([tuple(map(float, v)) for v in poly])
    o.processObstacles()
    o.set_pref_velocities(pref)
    o.doStep()
    sim.pref.copy_(torch.from_numpy(pref[None]))
    sim.doStep()
    gv, gp = sim.vel.cpu().numpy()[0], sim.pos.cpu().numpy()[0]
    ov, op = o.velocities(), o.positions()
    dv = np.abs(gv - ov).max(1)
    exact = float((dv == 0).mean())
    print(f"1M agents: worst dv={dv.max():.3g} worst dp={np.abs(gp - op).max():.3g} bit-exact agents={exact:.6f}")
    assert dv.max() <= TOL and np.abs(gp - op).max() <= TOL
    assert exact > 0.999          # the rest: ordered lists differing in bit-equal distances (tie exemption)


def test_collision_statistic_equals_the_oracle_count():
    """ORCA_STAT_COLLISIONS counts (agent, neighbor) pairs in RVO2's collision branch,
    distSq <= (r_i + r_j)^2, over the neighbor lists of the step -- compared with the same count
    taken from the oracle's lists (overlapping random crowds, tile and grid paths)."""
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    for scn in (scenarios.crowd(6, 64, seed=51), scenarios.crowd(1, 600, seed=52)):
        sims = oracle_sims(scn)
        gpu = _gpu_sim(scn)
        cr_sq = np.float32(2 * scn.params["radius"]) ** 2
        total_o = 0
        for t in range(12):
            pos = np.stack([s.positions() for s in sims])
            vel = np.stack([s.velocities() for s in sims])
            pref = goal_pref(pos, scn.goal).astype(np.float32)
            for e, s in enumerate(sims):
                s.set_pref_velocities(pref[e])
                s.doStep()
                for i in range(scn.agents_per_env):
                    total_o += sum(1 for _, d in s.agent_neighbors(i) if np.float32(d) <= cr_sq)
            gpu.pos.copy_(torch.from_numpy(pos))
            gpu.vel.copy_(torch.from_numpy(vel))
            gpu.pref.copy_(torch.from_numpy(pref))
            gpu.env_step(policy=_lib.POLICY_EXTERNAL)
        st = gpu.read_stats()
        assert total_o > 50
        assert st["collisions"] == total_o, (scn.name, st, total_o)
        assert st["agent_steps"] == 12 * scn.num_envs * scn.agents_per_env


def test_uncapped_obstacle_path_on_gpu():
    """Agents inside a ring of pillars see more than 16 obstacle edges / build more than 6 obstacle
    lines: the kernel redoes them in agent_slow_path (local-memory capacity 64) instead of dropping
    constraints.  Bit-identical to the oracle (which, like RVO2, has no cap); overflow stays 0."""
    import torch
    from _common import pillar_hall
    from collision_avoidance_b200 import scenarios
    worlds = [pillar_hall(seed=s) for s in range(24)]
    scn = scenarios.Scenario("pillar_hall", np.concatenate([w.pos for w in worlds]), np.concatenate([w.vel for w in worlds]),
                             np.concatenate([w.goal for w in worlds]), np.concatenate([w.goal2 for w in worlds]), 12.0,
                             worlds[0].obstacles)
    sims = oracle_sims(scn)
    gpu = _gpu_sim(scn)
    most_nbrs = most_lines = 0
    for t in range(200):
        pos = np.stack([s.positions() for s in sims])
        vel = np.stack([s.velocities() for s in sims])
        pref = goal_pref(pos, scn.goal).astype(np.float32)
        for e, s in enumerate(sims):
            s.set_pref_velocities(pref[e])
            s.doStep()
        gpu.pos.copy_(torch.from_numpy(pos))
        gpu.vel.copy_(torch.from_numpy(vel))
        gpu.pref.copy_(torch.from_numpy(pref))
        gpu.env_step(policy=0)
        assert np.array_equal(gpu.vel.cpu().numpy(), np.stack([s.velocities() for s in sims])), t
        assert np.array_equal(gpu.pos.cpu().numpy(), np.stack([s.positions() for s in sims])), t
        if t % 20 == 0:
            for s in sims:
                for i in range(scn.agents_per_env):
                    most_nbrs = max(most_nbrs, len(s.obstacle_neighbors(i)))
                    most_lines = max(most_lines, s.orca_lines(i)[1])
    assert most_nbrs > 16 and most_lines > 6, (most_nbrs, most_lines)
    assert gpu.read_stats()["overflow"] == 0


def test_every_shipped_scenario_finishes_without_overflow():
    """no scenario of the reference may ever drop an obstacle constraint."""
    import torch
    from collision_avoidance_b200 import alan
    for name, n in (("circle", 16), ("crowd", 24), ("blocks", 10), ("congested", 20), ("deadlock", 16), ("incoming", 17)):
        s = alan.Collision_Avoidance_Sim(numAgents=n, scenario=name, num_envs=8, seed=3)
        s.run_sim(mode=0, max_steps=1500)
        assert s.sim.read_stats()["overflow"] == 0, name
