"""GPU edge cases of the C-ABI path: ragged batch shapes, degenerate parameters, the env
policies on the uniform-grid path, error surfacing."""
import numpy as np
import pytest

from _common import goal_pref, oracle_sims

pytestmark = pytest.mark.gpu


def _mk(scn, **env):
    import torch
    from collision_avoidance_b200.sim import BatchedRVOSimulator
    sim = BatchedRVOSimulator(scn.num_envs, scn.agents_per_env, device="cuda:0", **scn.params)
    sim.set_obstacles(scn.obstacles, per_env=scn.per_env_obstacles)
    sim.pos.copy_(torch.from_numpy(scn.pos))
    sim.vel.copy_(torch.from_numpy(scn.vel))
    return sim


def _one_step_vs_oracle(scn, steps=20):
    import torch
    gpu = _mk(scn)
    sims = oracle_sims(scn)
    for _ in range(steps):
        pos = np.stack([s.positions() for s in sims])
        vel = np.stack([s.velocities() for s in sims])
        pref = goal_pref(pos, scn.goal).astype(np.float32)
        for e, s in enumerate(sims):
            s.set_pref_velocities(pref[e])
            s.doStep()
        gpu.pos.copy_(torch.from_numpy(pos))
        gpu.vel.copy_(torch.from_numpy(vel))
        gpu.pref.copy_(torch.from_numpy(pref))
        gpu.doStep()
        assert np.abs(gpu.vel.cpu().numpy() - np.stack([s.velocities() for s in sims])).max() <= 1e-4
        assert np.abs(gpu.pos.cpu().numpy() - np.stack([s.positions() for s in sims])).max() <= 1e-4


@pytest.mark.parametrize("E,N", [(1, 1), (3, 7), (5, 10), (37, 16), (2, 100), (3, 129), (1, 256), (2, 257), (1, 1000)])
def test_ragged_shapes_match_oracle(E, N):
    from collision_avoidance_b200 import scenarios
    _one_step_vs_oracle(scenarios.crowd(E, N, seed=E * 1000 + N), steps=8)


@pytest.mark.parametrize("k,nd", [(0, 5.0), (1, 5.0), (3, 2.0), (7, 5.0), (13, 3.0), (16, 8.0)])
def test_generic_max_neighbors(k, nd):
    from collision_avoidance_b200 import scenarios
    scn = scenarios.crowd(3, 40, seed=k + 50)
    scn.params = dict(scn.params, maxNeighbors=k, neighborDist=nd)
    _one_step_vs_oracle(scn, steps=10)


@pytest.mark.parametrize("N", [2, 10, 13, 16])
@pytest.mark.parametrize("k", [1, 3, 5, 7, 10, 16])
def test_ranked_selection_small_worlds(N, k):
    """Worlds of <= 16 agents pick their neighbors by rank counting (TileSource::gather_ranked16),
    not by sorted insertion: dense little crowds where more agents are in range than k allows."""
    from collision_avoidance_b200 import scenarios
    scn = scenarios.crowd(6, N, seed=100 * N + k)
    scn.params = dict(scn.params, maxNeighbors=k, neighborDist=5.0)
    _one_step_vs_oracle(scn, steps=6)


def test_no_obstacles_and_per_env_obstacles():
    from collision_avoidance_b200 import scenarios
    scn = scenarios.crowd(4, 30, seed=60)
    scn.obstacles = []
    _one_step_vs_oracle(scn, steps=10)
    _one_step_vs_oracle(scenarios.blocks(5, 14, seed=61), steps=15)


def test_env_policies_on_grid_path_equal_tile_path(monkeypatch):
    """ALAN bandit + done test, and the RL policy, stepped through the uniform-grid pipeline
    give the same bits as the shared-memory tile path (env counters bumped by the side kernel)."""
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    from collision_avoidance_b200.alan import DEFAULT_ONLINE_ACTIONS, unit_actions
    scn = scenarios.circle(6, 48, seed=70)
    E, N = 6, 48
    dev = "cuda:0"
    acts = torch.from_numpy(unit_actions(DEFAULT_ONLINE_ACTIONS)).to(dev)
    sims = []
    for force in (False, True):
        if force:
            monkeypatch.setenv("ORCA_B200_GRID_MIN_AGENTS", "2")
        sims.append(_mk(scn))
    monkeypatch.delenv("ORCA_B200_GRID_MIN_AGENTS")

    def fresh():
        return dict(goal=torch.from_numpy(scn.goal).to(dev), goal2=torch.from_numpy(scn.goal2).to(dev),
                    agent_done=torch.zeros(E, N, dtype=torch.uint8, device=dev),
                    arrival_time=torch.zeros(E, N, device=dev), env_step=torch.zeros(E, dtype=torch.int32, device=dev),
                    env_done_cnt=torch.zeros(E, dtype=torch.int32, device=dev), reward=torch.zeros(E, N, device=dev))
    st = [fresh(), fresh()]
    w = [torch.zeros(E, N, 8, device=dev), torch.zeros(E, N, 8, device=dev)]
    ids = [torch.zeros(E, N, dtype=torch.uint8, device=dev), torch.zeros(E, N, dtype=torch.uint8, device=dev)]
    for t in range(130):  # crosses the 121-step weight reset
        for s, state, ww, ii in zip(sims, st, w, ids):
            s.env_step(policy=_lib.POLICY_ALAN, done_mode=_lib.DONE_GOAL_RADIUS, alan_weights=ww, alan_actions=acts,
                       alan_action_out=ii, rng_seed=99, **state)
    assert torch.equal(sims[0].pos, sims[1].pos) and torch.equal(w[0], w[1]) and torch.equal(ids[0], ids[1])
    assert torch.equal(st[0]["env_step"], st[1]["env_step"]) and int(st[0]["env_step"][0]) == 130
    assert torch.equal(st[0]["reward"], st[1]["reward"])
    theta = (torch.rand(E, N, device=dev) - 0.5) * 2.0
    for s, state in zip(sims, st):
        s.env_step(policy=_lib.POLICY_RL, done_mode=_lib.DONE_X_BELOW, action_theta=theta, goal=state["goal"],
                   goal2=state["goal2"], agent_done=state["agent_done"], env_step=state["env_step"],
                   env_done_cnt=state["env_done_cnt"], reward=state["reward"])
    assert torch.equal(sims[0].vel, sims[1].vel) and torch.equal(st[0]["reward"], st[1]["reward"])


def test_errors_surface_as_python_exceptions():
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    from collision_avoidance_b200.sim import BatchedRVOSimulator
    with pytest.raises(NotImplementedError):
        BatchedRVOSimulator(2, 8, 1 / 60., 5.0, 17, 1.5, 1.5, 0.5, 1.0)       # maxNeighbors > 16
    with pytest.raises(ValueError):
        BatchedRVOSimulator(2, 8, 1 / 60., 5.0, 10, 1.5, 1.5, 0.5, 1.0, device="cpu")   # no CPU path
    sim = BatchedRVOSimulator(2, 8, 1 / 60., 5.0, 10, 1.5, 1.5, 0.5, 1.0)
    with pytest.raises(ValueError):
        sim.env_step(policy=_lib.POLICY_GOAL)                                   # goal missing
    with pytest.raises(ValueError):
        sim.env_step(policy=_lib.POLICY_GOAL, goal=torch.zeros(2, 8, 2))        # CPU tensor
    with pytest.raises(ValueError):
        sim.set_obstacles([[(0.0, 0.0)]])                                        # 1-vertex polygon
    with pytest.raises(ValueError):
        sim.env_step(policy=_lib.POLICY_GOAL, goal=torch.zeros(2, 8, 2, device="cuda"), done_mode=_lib.DONE_GOAL_RADIUS)


def test_host_buffer_entry_point_matches_device_path():
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    scn = scenarios.circle(64, 16, seed=80)
    a, b = _mk(scn), _mk(scn)
    goal = torch.from_numpy(scn.goal).cuda()
    pos_h = torch.from_numpy(scn.pos.copy()).pin_memory()
    vel_h = torch.from_numpy(scn.vel.copy()).pin_memory()
    goal_h = torch.from_numpy(scn.goal.copy()).pin_memory()
    b.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=True, steps=1)
    a.env_step(policy=_lib.POLICY_GOAL, goal=goal)
    for _ in range(20):
        a.env_step(policy=_lib.POLICY_GOAL, goal=goal)
        b.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
    assert np.array_equal(a.pos.cpu().numpy(), pos_h.numpy().reshape(64, 16, 2))
    assert np.array_equal(a.vel.cpu().numpy(), vel_h.numpy().reshape(64, 16, 2))


@pytest.mark.parametrize("mode", ["mapped", "pageable", "staged_graph", "staged_calls"])
@pytest.mark.parametrize("chunks", [2, 5, 16])
@pytest.mark.parametrize("world", ["circle", "crowd_blocks"])
def test_host_entry_point_pipelined_chunks_equal_single_launch(monkeypatch, chunks, world, mode):
    """orca_step_host has two routes.  Pinned (mapped) host buffers: the kernel reads the goals
    from and writes the new state into the caller's buffers itself.  Otherwise (pageable buffers, or
    ORCA_B200_HOST_NO_MAPPED): the batch is cut into env chunks on separate streams (upload | step |
    download overlapped), replayed from a CUDA graph or issued call by call.  Every route and every
    chunking must give the state of the plain device path, also with per-env obstacle tables and
    a ragged env count."""
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    if world == "circle":
        scn = scenarios.circle(67, 16, seed=81)
    else:
        scn = scenarios.crowd(13, 40, seed=82, blocks=4)   # per-env obstacle worlds
    E, N = scn.num_envs, scn.agents_per_env
    a, b = _mk(scn), _mk(scn)
    goal = torch.from_numpy(scn.goal).cuda()
    pos_h, vel_h, goal_h = (torch.from_numpy(x.copy()) for x in (scn.pos, scn.vel, scn.goal))
    if mode != "pageable":
        pos_h, vel_h, goal_h = pos_h.pin_memory(), vel_h.pin_memory(), goal_h.pin_memory()
    monkeypatch.setenv("ORCA_B200_HOST_CHUNKS", str(chunks))
    if mode.startswith("staged"):
        monkeypatch.setenv("ORCA_B200_HOST_NO_MAPPED", "1")
    if mode == "staged_calls":
        monkeypatch.setenv("ORCA_B200_HOST_NO_GRAPH", "1")   # call-by-call issue instead of graph replay
    b.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=True, steps=3)
    for _ in range(3):
        a.env_step(policy=_lib.POLICY_GOAL, goal=goal)
    for _ in range(10):
        a.env_step(policy=_lib.POLICY_GOAL, goal=goal)
        b.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
    assert np.array_equal(a.pos.cpu().numpy(), pos_h.numpy().reshape(E, N, 2))
    assert np.array_equal(a.vel.cpu().numpy(), vel_h.numpy().reshape(E, N, 2))
    # new obstacles invalidate the captured graph (its kernels carry the old table pointers)
    b.set_obstacles([], per_env=False)
    a.set_obstacles([], per_env=False)
    a.env_step(policy=_lib.POLICY_GOAL, goal=goal)
    b.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
    assert np.array_equal(a.pos.cpu().numpy(), pos_h.numpy().reshape(E, N, 2))


@pytest.mark.parametrize("E,N,blocks", [(9, 33, 0), (11, 48, 0), (3, 200, 0), (4, 256, 4), (5, 100, 4)])
def test_in_block_grid_equals_id_scan_on_gpu(monkeypatch, E, N, blocks):
    """Tile kernel, worlds of more than 32 agents: candidates from the in-block uniform grid
    (default) vs the plain scan over every agent of the env (ORCA_B200_NO_TILE_GRID) -- same bits
    after a long run, several envs per block included."""
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    scn = scenarios.crowd(E, N, seed=7 * N + E, blocks=blocks)
    goal = torch.from_numpy(scn.goal).cuda()
    a = _mk(scn)
    monkeypatch.setenv("ORCA_B200_NO_TILE_GRID", "1")
    b = _mk(scn)
    for t in range(150):
        monkeypatch.delenv("ORCA_B200_NO_TILE_GRID", raising=False)
        a.env_step(policy=_lib.POLICY_GOAL, goal=goal, want_neighbors=(t % 50 == 0))
        monkeypatch.setenv("ORCA_B200_NO_TILE_GRID", "1")
        b.env_step(policy=_lib.POLICY_GOAL, goal=goal, want_neighbors=(t % 50 == 0))
        if t % 50 == 0:
            assert torch.equal(a.nbr_idx, b.nbr_idx) and torch.equal(a.nbr_cnt, b.nbr_cnt)
    assert torch.equal(a.pos, b.pos) and torch.equal(a.vel, b.vel)
    assert a.read_stats() == b.read_stats()


@pytest.mark.parametrize("R,C", [(16, 8), (12, 6), (5, 3), (32, 16), (1, 8), (9, 5)])
def test_observation_kernel_equals_its_host_twin(R, C):
    """observe_kernel (block-staged, queue of (ray, neighbor) pairs) against the serial host twin
    of the same functions (tests/host_emul), for ray / polygon counts other than the default 16 / 8:
    whole agents per block (21 for 12 rays, 32 for 5 ...), ragged last block, per-env obstacles,
    k = 10 neighbor lists.  Same float32 operations in the same order -> identical bits."""
    import torch
    import _emul
    from _common import snake
    from collision_avoidance_b200 import _lib, scenarios
    scn = scenarios.crowd(7, 23, seed=31 + R, blocks=4)
    scn.params = dict(scn.params, neighborDist=3.0)
    sim = _mk(scn)
    goal = torch.from_numpy(scn.goal).cuda()
    for _ in range(25):
        sim.env_step(policy=_lib.POLICY_GOAL, goal=goal, want_neighbors=True)
    obs = sim.observe(goal, laser_num=R, circle_approx_num=C).cpu().numpy()
    pos, vel = sim.pos.cpu().numpy(), sim.vel.cpu().numpy()
    P = snake(scn.params)
    hits = 0
    for e in range(scn.num_envs):
        nbr = {"nbr_idx": np.ascontiguousarray(sim.nbr_idx[e:e + 1].cpu().numpy()),
               "nbr_cnt": np.ascontiguousarray(sim.nbr_cnt[e:e + 1].cpu().numpy()),
               "onbr_idx": np.ascontiguousarray(sim.obst_nbr_idx[e:e + 1].cpu().numpy()),
               "onbr_cnt": np.ascontiguousarray(sim.obst_nbr_cnt[e:e + 1].cpu().numpy())}
        ref = _emul.emul_observe(P, pos[e:e + 1].copy(), vel[e:e + 1].copy(), scn.goal[e:e + 1].copy(), nbr,
                                 world=_emul.World(scn.obstacles[e]), laser_num=R, circle_approx_num=C)
        assert np.array_equal(ref, obs[e:e + 1]), (e, float(np.abs(ref - obs[e:e + 1]).max()))
        hits += int((np.abs(ref).reshape(-1, 4)[:, :2].sum(-1) > 0).sum())
    assert hits > 0 or R == 1


@pytest.mark.parametrize("E", [1, 23, 64, 320])
def test_observation_of_the_gym_world_equals_its_host_twin(E):
    """The gym world (10 agents, k = 5, radius 0.5 of 1.5: large polygons, many (ray, neighbor) pairs,
    shared obstacles) at batch sizes whose last warp chunk is partial -- 640 and 3,200 agents are the
    sizes at which a ptxas miscompile of the chunk bound once ran past the end of the batch (see
    observe_kernel) -- against the serial host twin, bit for bit, walls and blocks included."""
    import torch
    import _emul
    from _common import snake
    from collision_avoidance_b200 import envs
    env = envs.Collision_Avoidance_Env(numAgents=10, num_envs=E, seed=3)
    theta = (torch.rand(E, 10, device="cuda", generator=torch.Generator("cuda").manual_seed(5)) - 0.5) * 0.6
    for _ in range(40):
        env.step(theta)
    obs = env.obs.cpu().numpy()
    pos, vel, goal = env.sim.pos.cpu().numpy(), env.sim.vel.cpu().numpy(), env.targets_pos.cpu().numpy()
    nbr = {"nbr_idx": np.ascontiguousarray(env.sim.nbr_idx.cpu().numpy()),
           "nbr_cnt": np.ascontiguousarray(env.sim.nbr_cnt.cpu().numpy()),
           "onbr_idx": np.ascontiguousarray(env.sim.obst_nbr_idx.cpu().numpy()),
           "onbr_cnt": np.ascontiguousarray(env.sim.obst_nbr_cnt.cpu().numpy())}
    ref = _emul.emul_observe(snake(env.scn.params), pos, vel, goal, nbr, world=_emul.World(env.scn.obstacles))
    assert np.array_equal(ref, obs), float(np.abs(ref - obs).max())
    assert (np.abs(ref).reshape(-1, 4)[:, :2].sum(-1) > 0).sum() > E  # rays do hit things
    assert nbr["onbr_cnt"].max() > 0


def test_handles_on_two_devices_in_one_process():
    """Kernel attributes (dynamic shared memory opt-in) are per device: a second handle on another
    GPU of the same process must launch just like the first.  Skipped on single-GPU boxes."""
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    from collision_avoidance_b200.sim import BatchedRVOSimulator
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    scn = scenarios.circle(40, 16, seed=5)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        sim = BatchedRVOSimulator(scn.num_envs, scn.agents_per_env, device=dev, **scn.params)
        sim.set_obstacles(scn.obstacles)
        sim.pos.copy_(torch.from_numpy(scn.pos))
        sim.vel.copy_(torch.from_numpy(scn.vel))
        goal = torch.from_numpy(scn.goal).to(dev)
        for _ in range(30):
            sim.env_step(policy=_lib.POLICY_GOAL, goal=goal)
        outs.append(sim.pos.cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("mode", ["mapped", "staged"])
def test_host_entry_point_reduced_traffic_variants(monkeypatch, mode):
    """orca_step_host_ex: positions-only write-back (vel_host = NULL) and ORCA_HOST_AUX_UNCHANGED (the
    goal buffer is not re-read) give the bits of the full-traffic call on both routes; a changed goal
    buffer is picked up again as soon as the flag is dropped; stepping before any state upload is
    ORCA_ERR_STATE, not garbage."""
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    if mode == "staged":
        monkeypatch.setenv("ORCA_B200_HOST_NO_MAPPED", "1")
    scn = scenarios.circle(70, 16, seed=91)
    E, N = scn.num_envs, scn.agents_per_env
    full, lean = _mk(scn), _mk(scn)
    pin = lambda x: torch.from_numpy(x.copy()).pin_memory()   # noqa: E731
    pf, vf, gf = pin(scn.pos), pin(scn.vel), pin(scn.goal)
    pl, vl, gl = pin(scn.pos), pin(scn.vel), pin(scn.goal)
    with pytest.raises(RuntimeError):
        lean.step_host(pl, None, gl, policy=_lib.POLICY_GOAL, upload_state=False)
    full.step_host(pf, vf, gf, policy=_lib.POLICY_GOAL, upload_state=True)
    lean.step_host(pl, vl, gl, policy=_lib.POLICY_GOAL, upload_state=True, aux_unchanged=True)
    for t in range(12):
        full.step_host(pf, vf, gf, policy=_lib.POLICY_GOAL, upload_state=False)
        lean.step_host(pl, None, gl, policy=_lib.POLICY_GOAL, upload_state=False, aux_unchanged=True)
        assert torch.equal(pf, pl), t
    assert not torch.equal(vf, vl)                  # the lean call never wrote velocities back ...
    # ... but its device state carries them: a full call afterwards returns the same velocities
    gf.copy_(torch.from_numpy(scn.pos))             # new goals (everybody walks home): flag dropped -> re-read
    gl.copy_(torch.from_numpy(scn.pos))
    full.step_host(pf, vf, gf, policy=_lib.POLICY_GOAL, upload_state=False)
    lean.step_host(pl, vl, gl, policy=_lib.POLICY_GOAL, upload_state=False)
    assert torch.equal(pf, pl) and torch.equal(vf, vl)


@pytest.mark.parametrize("N", [8, 300])
def test_finished_worlds_stop_counting_statistics(N):
    """Worlds that finish early keep being stepped (run_sim polls rarely), but their idle steps do not
    count into the episode statistics -- tile path and uniform-grid path."""
    import torch
    from collision_avoidance_b200 import _lib, scenarios
    scn = scenarios.crowd(2, N, seed=3)
    scn.goal[0] = scn.pos[0]            # world 0: everybody already stands on the goal -> done after step 1
    scn.goal2[0] = scn.pos[0]
    sim = _mk(scn)
    E = 2
    st = dict(goal=torch.from_numpy(scn.goal).cuda(), goal2=torch.from_numpy(scn.goal2).cuda(),
              agent_done=torch.zeros(E, N, dtype=torch.uint8, device="cuda"), arrival_time=torch.zeros(E, N, device="cuda"),
              env_step=torch.zeros(E, dtype=torch.int32, device="cuda"), env_done_cnt=torch.zeros(E, dtype=torch.int32, device="cuda"))
    steps = 20
    for _ in range(steps):
        sim.env_step(policy=_lib.POLICY_GOAL, done_mode=_lib.DONE_GOAL_RADIUS, **st)
    assert int(st["env_done_cnt"][0]) == N and int(st["env_done_cnt"][1]) < N
    assert st["env_step"].tolist() == [steps, steps]                 # both worlds were stepped ...
    s = sim.read_stats()
    assert s["agent_steps"] == N + steps * N                         # ... world 0 counted for its one live step only
    assert s["finished"] >= N
