"""CUDA env / ALAN / trainer paths against fixtures recorded from the UNMODIFIED reference shells
(tests/golden/shell_*.npz, made by tests/golden/make_shell_golden.py; CPU-side pins of the same
fixtures: tests/test_shell_golden.py).

Method: "single steps from shared states" (BASELINE north_star), batched -- every recorded
frame t of a reference run becomes one WORLD of a GPU batch, all worlds take one fused step in
one launch, and world t must land on the reference's frame t + 1.

Tolerances (the contract): positions / velocities 1e-4 absolute (north star); rewards, bandit
weights, arrival times 1e-5 (float32 kernel vs the reference's float64 shell); action ids equal
except draws within float32 rounding of a CDF boundary; observation 2e-4 on rays where both
sides agree on hit-or-miss, at most 0.2 % of rays may flip (grazing rays)."""
import numpy as np
import pytest

from _golden import alan_fixture_names, load_alan, load_env

pytestmark = pytest.mark.gpu

TOL_STATE = 1e-4
TOL_SHELL = 1e-5


def _alan_batch(fx, frames, mode, copies=1):
    """GPU ALAN shell whose world w holds frame frames[w % len(frames)] of the fixture."""
    import torch
    from collision_avoidance_b200 import alan
    N = int(fx["numAgents"])
    frames = np.asarray(frames)
    idx = np.tile(frames, copies)
    E = len(idx)
    actions = [tuple(a) for a in fx["actions"]]
    gpu = alan.Collision_Avoidance_Sim(numAgents=N, scenario="circle", online_actions=actions, num_envs=E, seed=1)
    gpu.sim.set_obstacles([[tuple(map(float, v)) for v in poly] for poly in fx["polygons"]])
    dev = gpu.device
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else a.astype(dt))).to(dev)   # noqa: E731
    gpu.sim.pos.copy_(t(fx["pos"][idx]))
    gpu.sim.vel.copy_(t(fx["vel"][idx]))
    goal = fx["tgt"][idx].astype(np.float32)
    done = fx["done"][idx].copy()
    if mode == 0:
        # run_sim(mode=0): an agent that arrived in step t still aims at its OLD target in step t + 1
        # (update_pref_vel ran before done_test swapped it) -> "swap pending" flag 2 + the old goal
        prev = np.maximum(idx - 1, 0)
        fresh = (fx["done"][idx] == 1) & (fx["done"][prev] == 0) & (idx > 0)[:, None]
        goal[fresh] = fx["tgt"][prev].astype(np.float32)[fresh]
        done[fresh] = 2
    gpu.goal.copy_(t(goal))
    gpu.goal2.copy_(t(np.broadcast_to(fx["tgt2"].astype(np.float32), goal.shape)))
    gpu.agents_done.copy_(t(done))
    gpu.env_done_cnt.copy_(t((done != 0).sum(1), np.int32))
    gpu.env_step.copy_(t(idx, np.int32))
    gpu.agents_time.copy_(t(fx["atime"][idx], np.float32))
    if mode == 1:
        gpu.action_weights.copy_(t(fx["w"][idx], np.float32))
    return gpu, idx


@pytest.mark.parametrize("name", alan_fixture_names(mode=1))
def test_cuda_online_step_lands_on_the_reference_frames(name):
    """ALAN_true.py:569-628,547-566 -- incl. circle N = 32 (BASELINE config 3's world shape),
    `incoming`, two-stage goals (deadlock, congested), per-env obstacle blocks."""
    import torch
    fx = load_alan(name)
    T = fx["pos"].shape[0] - 1
    gpu, idx = _alan_batch(fx, np.arange(T), mode=1)
    gpu.online_step(uniforms=torch.from_numpy(fx["u"][idx].astype(np.float32)).to(gpu.device))
    act = gpu.action_ids.cpu().numpy()
    same = act == fx["aid"][idx]
    assert (~same).sum() <= 2, f"{(~same).sum()} action ids differ"
    nxt = idx + 1
    gp, gv = gpu.sim.pos.cpu().numpy(), gpu.sim.vel.cpu().numpy()
    worst = max(np.abs(gp - fx["pos"][nxt])[same].max(), np.abs(gv - fx["vel"][nxt])[same].max())
    exact = float((gv == fx["vel"][nxt])[same].mean())
    w = gpu.action_weights.cpu().numpy()
    worst_w = np.abs(w - fx["w"][nxt])[same].max()
    chosen = np.take_along_axis(fx["w"][nxt], fx["aid"][idx][..., None].astype(np.int64), -1)[..., 0]
    worst_r = np.abs(gpu.reward.cpu().numpy() - chosen)[same].max()      # weight[a] = R (:628)
    print(f"{name}: state={worst:.3g} (bit-equal velocities {exact:.3f}) weights={worst_w:.3g} reward={worst_r:.3g} "
          f"action mismatches={(~same).sum()}/{same.size}")
    assert worst <= TOL_STATE and worst_w <= TOL_SHELL and worst_r <= TOL_SHELL
    d = gpu.agents_done.cpu().numpy()
    assert np.array_equal(d[same.all(1)], fx["done"][nxt][same.all(1)])
    arrived = (fx["done"][nxt] == 1) & same
    if arrived.any():
        assert np.abs(gpu.agents_time.cpu().numpy() - fx["atime"][nxt])[arrived].max() <= TOL_SHELL
    assert np.array_equal(gpu.goal.cpu().numpy()[same.all(1)], fx["tgt"][nxt].astype(np.float32)[same.all(1)])
    assert np.array_equal(gpu.env_step.cpu().numpy(), nxt)


@pytest.mark.parametrize("name", alan_fixture_names(mode=0))
def test_cuda_orca_step_lands_on_the_reference_frames(name):
    """ALAN_true.py:631-636 under run_sim(mode=0), incl. the one-step-late goal swap."""
    fx = load_alan(name)
    T = fx["pos"].shape[0] - 1
    gpu, idx = _alan_batch(fx, np.arange(T), mode=0)
    gpu.orca_step()
    nxt = idx + 1
    gp, gv = gpu.sim.pos.cpu().numpy(), gpu.sim.vel.cpu().numpy()
    # An agent parked ON its goal has an ill-conditioned goal direction: the reference keeps the goal
    # in float64, the C ABI takes float32 goals (rounding <= 1e-6 at these coordinates), and the unit
    # vector to a point d away turns by up to 1e-6 / d.  Within d < 0.02 that exceeds the 1e-4 bar
    # for reasons outside the step, so those agents are only required to respect the speed limit.
    d_goal = np.linalg.norm(fx["tgt"][idx] - fx["pos"][idx].astype(np.float64), axis=-1)
    ok = d_goal >= 0.02
    assert ok.mean() > 0.8
    worst = max(np.abs(gp - fx["pos"][nxt])[ok].max(), np.abs(gv - fx["vel"][nxt])[ok].max())
    print(f"{name}: state={worst:.3g} bit-equal velocities {(gv == fx['vel'][nxt])[ok].mean():.3f} "
          f"agents parked on their goal {(~ok).sum()}/{ok.size}")
    assert worst <= TOL_STATE
    assert np.linalg.norm(gv[~ok], axis=-1).max(initial=0.0) <= 1.0 + 1e-5
    assert np.array_equal(gpu.agents_done.cpu().numpy() != 0, fx["done"][nxt] != 0)
    arrived = fx["done"][nxt] == 1
    if arrived.any():
        assert np.abs(gpu.agents_time.cpu().numpy() - fx["atime"][nxt])[arrived].max() <= TOL_SHELL
    # goals: swapped for agents that arrived before this step, still the old one for fresh arrivals
    fresh = (fx["done"][nxt] == 1) & (fx["done"][idx] == 0)
    g = gpu.goal.cpu().numpy()
    assert np.array_equal(g[~fresh], fx["tgt"][nxt].astype(np.float32)[~fresh])
    assert np.array_equal(g[fresh], fx["tgt"][idx].astype(np.float32)[fresh])
    assert (gpu.agents_done.cpu().numpy()[fresh] == 2).all()


def test_cuda_orca_run_follows_the_reference_run_closed_loop():
    """Closed loop (no re-synchronisation): the CUDA orca_step walks the recorded reference run of
    the 24-agent crowd for as long as float32-vs-float64 goal directions keep the (chaotic)
    trajectories together."""
    fx = load_alan("shell_orca_crowd24")
    gpu, idx = _alan_batch(fx, [0], mode=0)
    T = fx["pos"].shape[0] - 1
    worst = []
    for t in range(T):
        gpu.orca_step()
        worst.append(float(np.abs(gpu.sim.pos.cpu().numpy()[0] - fx["pos"][t + 1]).max()))
    print(f"closed loop: max |dpos| after 10/50/{T} steps = {worst[9]:.2g} / {worst[49]:.2g} / {worst[-1]:.2g}")
    assert worst[9] <= 1e-4


def test_cuda_alan_full_size_batch_of_reference_frames():
    """BASELINE config 3's per-GPU share at full size (32,768 worlds x 32 agents, 8 actions): the
    batch is filled with the recorded frames, every copy must land on its reference frame."""
    import torch
    fx = load_alan("shell_alan_circle32")
    T = fx["pos"].shape[0] - 1
    frames = np.arange(T - T % 32)
    copies = 32768 // len(frames)
    frames = np.arange(32768 // copies)
    gpu, idx = _alan_batch(fx, frames, mode=1, copies=copies)
    assert gpu.num_envs * gpu.numAgents >= 1_000_000
    gpu.online_step(uniforms=torch.from_numpy(fx["u"][idx].astype(np.float32)).to(gpu.device))
    same = gpu.action_ids.cpu().numpy() == fx["aid"][idx]
    assert (~same).sum() <= 2 * copies
    gv = gpu.sim.vel.cpu().numpy()
    assert np.abs(gv - fx["vel"][idx + 1])[same].max() <= TOL_STATE
    blocks = gv.reshape(copies, len(frames), gpu.numAgents, 2)
    assert np.array_equal(blocks, np.broadcast_to(blocks[0], blocks.shape))      # copies are bit-identical


def test_cuda_run_sim_ttime_formula_equals_the_reference():
    """ALAN_true.py:125-131: TTime = mean + 3 sigma of the per-agent arrival times."""
    import torch
    for name in ("shell_alan_circle16", "shell_orca_congested12", "shell_alan_blocks8"):
        fx = load_alan(name)
        gpu, _ = _alan_batch(fx, [0], mode=int(fx["mode"]))
        gpu.agents_time.copy_(torch.from_numpy(fx["final_atime"].astype(np.float32))[None])
        gpu.agents_done.copy_(torch.from_numpy(fx["final_done"])[None])
        gpu.env_done_cnt.fill_(int(fx["final_done"].sum()))
        success, total_time, ttime, _ = gpu.run_sim(mode=int(fx["mode"]), max_steps=0)
        assert bool(success[0]) == bool(fx["result"][0])
        assert abs(float(ttime[0]) - fx["result"][2]) <= 1e-5 * fx["result"][2]


def test_cuda_reference_rng_worlds_start_like_the_reference():
    """alan.Collision_Avoidance_Sim(reference_rng=True): world e == the reference after
    random.seed(seed + e): same start state, same min TTime (ALAN_true.py:161-172)."""
    import json
    import os
    from collision_avoidance_b200 import alan
    from _golden import GOLDEN
    with open(os.path.join(GOLDEN, "shell_scenarios.json")) as f:
        recs = json.load(f)["alan"]
    for rec in recs:
        gpu = alan.Collision_Avoidance_Sim(numAgents=rec["numAgents"], scenario=rec["name"], num_envs=2, seed=rec["seed"] - 1,
                                           reference_rng=True)
        assert np.array_equal(gpu.sim.pos.cpu().numpy()[1], np.asarray(rec["pos"], np.float32)), rec["name"]
        assert np.array_equal(gpu.sim.vel.cpu().numpy()[1], np.asarray(rec["vel"], np.float32)), rec["name"]
        assert abs(float(gpu.min_TTime[1]) - rec["min_TTime_after_reset"]) <= 1e-5 * rec["min_TTime_after_reset"]
        assert gpu.max_step == rec["max_step"]


# ------------------------------------------------------------------------------------ gym env
def _obs_compare(g_obs, o_obs):
    g = g_obs.reshape(-1, 16, 4)
    o = o_obs.reshape(-1, 16, 4)
    hit_g = np.abs(g[..., :2]).sum(-1) > 0
    hit_o = np.abs(o[..., :2]).sum(-1) > 0
    agree = hit_g == hit_o
    worst = float(np.abs(g - o)[agree].max()) if agree.any() else 0.0
    return worst, int((~agree).sum()), agree.size


@pytest.mark.parametrize("name", ["shell_env", "shell_env_small"])
def test_cuda_env_step_lands_on_the_reference_rows(name):
    """collision_avoidence_env.py:367-416 (step), :447-458 (orca_step), :231-318 (_get_obs): every
    recorded transition is one world of a batch."""
    import torch
    from collision_avoidance_b200 import _lib, envs
    fx = load_env(name)
    N = int(fx["numAgents"])
    kinds = fx["kind"]
    for kind in (1, 3):
        rows = np.where(kinds == kind)[0]
        if len(rows) == 0:
            continue
        prev = rows - 1
        E = len(rows)
        env = envs.Collision_Avoidance_Env(numAgents=N, num_envs=E, seed=3)
        dev = env.device
        t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else a.astype(dt))).to(dev)   # noqa: E731
        env.sim.pos.copy_(t(fx["pos"][prev]))
        env.sim.vel.copy_(t(fx["vel"][prev]))
        env.targets_pos.copy_(t(fx["tgt"][prev], np.float32))
        env.agents_done.copy_(t(fx["agents_done"][prev]))
        env.env_done_cnt.copy_(t(fx["agents_done"][prev].sum(1), np.int32))
        env.env_step.copy_(t(fx["step_count"][prev], np.int32))
        if kind == 1:
            obs, rew, done, _ = env.step(t(fx["theta"][rows], np.float32))
            assert np.abs(rew.cpu().numpy() - fx["rew"][rows]).max() <= TOL_SHELL
            assert np.array_equal(done.cpu().numpy(), fx["done"][rows].astype(bool))
            assert np.array_equal(env.agents_done.cpu().numpy(), fx["agents_done"][rows])
            assert np.array_equal(env.targets_pos.cpu().numpy(), fx["tgt"][rows].astype(np.float32))
        else:
            env.orca_step()
        gp, gv = env.sim.pos.cpu().numpy(), env.sim.vel.cpu().numpy()
        worst = max(np.abs(gp - fx["pos"][rows]).max(), np.abs(gv - fx["vel"][rows]).max())
        assert worst <= TOL_STATE
        # neighbor lists of the step (pre-update positions) == the reference's getters
        assert np.array_equal(env.sim.nbr_cnt.cpu().numpy(), fx["na"][rows])
        assert np.array_equal(env.sim.nbr_idx.cpu().numpy(), fx["ia"][rows])
        assert np.array_equal(env.sim.obst_nbr_cnt.cpu().numpy(), fx["no"][rows])
        assert np.array_equal(env.sim.obst_nbr_idx.cpu().numpy(), fx["io"][rows])
        # observation from the reference's exact post-step state (the GPU's is within 1e-4 of it)
        env.sim.pos.copy_(t(fx["pos"][rows]))
        env.sim.vel.copy_(t(fx["vel"][rows]))
        w, flips, rays = _obs_compare(env._get_obs().cpu().numpy(), fx["obs"][rows])
        print(f"{name} kind {kind}: {E} transitions, state={worst:.3g} obs={w:.3g} hit/miss flips={flips}/{rays}")
        assert w <= 2e-4 and flips <= 0.002 * rays


def test_cuda_env_reset_keeps_stale_lists_like_the_reference():
    """collision_avoidence_env.py:461-488 + SURVEY Q3/Q4: reset() re-draws positions only; the
    observation it returns combines the NEW positions with the neighbor lists of the last doStep."""
    import torch
    from collision_avoidance_b200 import envs
    fx = load_env("shell_env")
    N = int(fx["numAgents"])
    rows = np.where(fx["kind"] == 2)[0]
    assert len(rows) >= 2
    for r in rows:
        env = envs.Collision_Avoidance_Env(numAgents=N, num_envs=1, seed=3)
        dev = env.device
        t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else a.astype(dt))).to(dev)   # noqa: E731
        # the world as the reference left it before reset(): state + lists of row r - 1
        env.sim.vel.copy_(t(fx["vel"][r - 1][None]))
        env.targets_pos.copy_(t(fx["tgt"][r - 1][None], np.float32))
        env.sim.nbr_cnt.copy_(t(fx["na"][r - 1][None]))
        env.sim.nbr_idx.copy_(t(fx["ia"][r - 1][None]))
        env.sim.obst_nbr_cnt.copy_(t(fx["no"][r - 1][None]))
        env.sim.obst_nbr_idx.copy_(t(fx["io"][r - 1][None]))
        env.reset()
        assert (env.sim.vel.cpu().numpy()[0] == fx["vel"][r - 1]).all()          # velocities survive
        assert int(env.agents_done.sum()) == 0 and int(env.env_step[0]) == 0
        env.sim.pos.copy_(t(fx["pos"][r][None]))                                  # the reference's draw
        w, flips, rays = _obs_compare(env._get_obs().cpu().numpy(), fx["obs"][r][None])
        assert w <= 2e-4 and flips <= 2, (r, w, flips)


def test_cuda_env_reference_rng_and_gym_surface():
    """reference_rng worlds == the reference's spawn / reset draws; spaces, registry id and
    reset(env_mask) as a vector env would use them (row f4)."""
    import json
    import os
    import torch
    from collision_avoidance_b200 import envs, spaces
    from _golden import GOLDEN
    with open(os.path.join(GOLDEN, "shell_scenarios.json")) as f:
        rec = json.load(f)["env"][0]
    env = spaces.make(spaces.ENV_ID, numAgents=rec["numAgents"], num_envs=3, seed=rec["seed"] - 2, reference_rng=True)
    assert isinstance(env, envs.Collision_Avoidance_Env)
    assert np.array_equal(env.sim.pos.cpu().numpy()[2], np.asarray(rec["pos"], np.float32))
    assert np.array_equal(env.sim.vel.cpu().numpy()[2], np.asarray(rec["vel"], np.float32))
    assert env.action_space == spaces.Box(-np.pi, np.pi, (1,)) and env.observation_space.shape == (64,)
    a = env.action_space.sample((3, rec["numAgents"]))[..., 0]
    assert env.action_space.contains(a[..., None])
    obs, rew, done, info = env.step(torch.from_numpy(a))
    assert obs.shape == (3, rec["numAgents"], 64) and float(obs.abs().max()) <= 1.5 + 1e-6
    before = env.sim.pos.clone()
    env.reset(env_mask=torch.tensor([False, True, False]))
    assert torch.equal(env.sim.pos[0], before[0]) and torch.equal(env.sim.pos[2], before[2])
    assert not torch.equal(env.sim.pos[1], before[1])
    assert int(env.env_step[1]) == 0 and int(env.env_step[0]) == 1
