"""CPU side of the reference-shell pins (no GPU needed).

tests/golden/shell_*.{npz,json,json.gz} were recorded by running the UNMODIFIED reference shells
(ALAN_true.py, collision_avoidence_env.py, Train_ALAN_action_space.py) under stubs with
``rvo2`` bound to the CPU oracle (tests/golden/make_shell_golden.py).  Here:

  * the float64 shell restatement (oracle/shell_oracle.py) must reproduce them EXACTLY -- this is
    what licenses it as the per-step checker of the CUDA paths in the ``-m gpu`` tests;
  * the host-side product logic that can run without a GPU -- scenario generators with
    ``reference_rng``, the MCMC trainer with ``reference_semantics``, the ALAN window period,
    the gym surface objects -- must reproduce them exactly as well;
  * the recorded rvo2-boundary call trace replays against the oracle's PyRVOSimulator (the same
    replay runs against ``rvo2_compat`` on the GPU in tests/test_gpu_compat.py).
"""
import json
import os

import numpy as np
import pytest

from _golden import (GOLDEN, alan_fixture_names, load_alan, load_calltrace, load_env, oracle_from_alan_fixture,
                     oracle_from_env_fixture, replay_calltrace, scenario_from_fixture)
from collision_avoidance_b200 import scenarios
from oracle import rvo2_oracle


# ------------------------------------------------------------------------------ scenarios (a16, f3)
def _scn_fixtures():
    with open(os.path.join(GOLDEN, "shell_scenarios.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("rec", _scn_fixtures()["alan"], ids=lambda r: "%s%d" % (r["name"], r["numAgents"]))
def test_scenario_generators_equal_the_reference_worlds(rec):
    """ALAN_true.py:175-457: all six ``_init_world_*`` (positions, initial velocities, two-stage
    targets, obstacle polygons, envsize) under the reference's own random stream."""
    kw = dict(rotate=False) if rec["name"] == "circle" else {}
    s = scenarios.make(rec["name"], 1, rec["numAgents"], seed=rec["seed"], reference_rng=True, **kw)
    t = np.asarray(rec["targets"])
    polys = s.obstacles[0] if s.per_env_obstacles else s.obstacles
    assert np.array_equal(s.pos[0], np.asarray(rec["pos"], np.float32))
    assert np.array_equal(s.vel[0], np.asarray(rec["vel"], np.float32))
    assert np.array_equal(s.goal[0], t[:, 0].astype(np.float32))
    assert np.array_equal(s.goal2[0], t[:, 1].astype(np.float32))
    assert np.array_equal(np.asarray(polys, np.float32), np.asarray(rec["polygons"], np.float32))
    assert s.envsize == rec["envsize"]
    assert rec["min_TTime_after_ctor"] == 0      # reference quirk: __init__ zeroes it after _init_world


def test_scenario_batches_are_independent_reference_worlds():
    """World e of a batch == the reference after random.seed(seed + e)."""
    rec = [r for r in _scn_fixtures()["alan"] if r["name"] == "crowd"][0]
    s = scenarios.make("crowd", 3, rec["numAgents"], seed=rec["seed"] - 2, reference_rng=True)
    assert np.array_equal(s.pos[2], np.asarray(rec["pos"], np.float32))
    assert not np.array_equal(s.pos[0], s.pos[2])


@pytest.mark.parametrize("rec", _scn_fixtures()["env"], ids=lambda r: "env%d" % r["numAgents"])
def test_default_env_world_and_gym_surface(rec):
    """collision_avoidence_env.py:52-53,77-123,321-350,461-488 + collision_avoidance/__init__.py:1-6."""
    from collision_avoidance_b200 import spaces
    streams = scenarios.reference_streams(rec["seed"], 1)
    s = scenarios.default_env(1, rec["numAgents"], reference_rng=streams)
    pos = scenarios.default_env_reset_positions(streams, rec["numAgents"])     # __init__ ends with reset()
    assert np.array_equal(pos[0], np.asarray(rec["pos"], np.float32))
    assert np.array_equal(s.vel[0], np.asarray(rec["vel"], np.float32))
    assert np.array_equal(s.goal[0], np.asarray(rec["targets"], np.float32))
    assert np.array_equal(np.asarray(s.obstacles, np.float32), np.asarray(rec["polygons"], np.float32))
    a, o = spaces.env_spaces(neighbor_dist=1.5, laser_num=16)
    assert [a.low, a.high, list(a.shape)] == rec["action_space"]
    assert [o.low, o.high, list(o.shape)] == rec["observation_space"]
    assert spaces.ENV_ID in rec["registry"] and spaces.ENV_ID in spaces.registry
    from oracle import shell_oracle
    assert np.allclose(np.asarray(shell_oracle.laser_rays(16, 1.5)), np.asarray(rec["ray_lines"]), atol=0, rtol=0)
    assert np.allclose(np.asarray(shell_oracle.circle_approx(8, 0.5)), np.asarray(rec["approx_lines"]), atol=0, rtol=0)


# ------------------------------------------------------------------------------ ALAN shell (a10-a14, a16)
@pytest.mark.parametrize("name", alan_fixture_names(mode=1))
def test_shell_oracle_online_step_equals_reference_run(name):
    """ALAN_true.py:569-628 + :547-566: the restated shell, fed the recorded uniforms, walks the
    reference's own trajectory: action ids, weights, arrival times, states -- exactly."""
    fx = load_alan(name)
    sh = oracle_from_alan_fixture(fx)
    T_rec = fx["pos"].shape[0] - 1
    for t in range(T_rec):
        sh.online_step(fx["u"][t])
        sh.step_count += 1
        sh.done_test()
        assert sh.last["action_ids"] == list(fx["aid"][t]), (name, t)
        assert np.array_equal(np.asarray(sh.action_weights), fx["w"][t + 1]), (name, t)
        assert np.array_equal(sh.sim.positions(), fx["pos"][t + 1]), (name, t)
        assert np.array_equal(sh.sim.velocities(), fx["vel"][t + 1]), (name, t)
        assert sh.agents_done == list(fx["done"][t + 1])
        assert np.array_equal(np.asarray(sh.agents_time), fx["atime"][t + 1])


@pytest.mark.parametrize("name", ["shell_alan_circle16", "shell_alan_crowd12", "shell_alan_incoming10"])
def test_shell_oracle_full_episode_ttime_equals_reference(name):
    """ALAN_true.py:106-131: whole ``run_sim(1)`` -> (success, total_time, TTime)."""
    fx = load_alan(name)
    sh = oracle_from_alan_fixture(fx)
    success, total_time, ttime, _ = sh.run_sim(mode=1, uniforms=fx["u"])
    assert sh.step_count == int(fx["steps"])
    assert np.array_equal(np.asarray(sh.agents_time), fx["final_atime"])
    assert [float(success), total_time, ttime] == list(fx["result"][:3])


@pytest.mark.parametrize("name", alan_fixture_names(mode=0))
def test_shell_oracle_orca_step_equals_reference_run(name):
    """ALAN_true.py:631-636 + run_sim(mode=0): states, preferred velocities (aimed at the
    PRE-swap target on the step after an arrival), arrival times, TTime."""
    fx = load_alan(name)
    sh = oracle_from_alan_fixture(fx)
    T_rec = fx["pos"].shape[0] - 1
    for t in range(T_rec):
        sh.orca_step()
        sh.step_count += 1
        sh.done_test()
        assert np.array_equal(sh.sim.positions(), fx["pos"][t + 1]), (name, t)
        pref = np.asarray([sh.sim.getAgentPrefVelocity(i) for i in range(sh.N)], np.float32)
        assert np.array_equal(pref, fx["pref"][t + 1]), (name, t)
    success, total_time, ttime, _ = sh.run_sim(mode=0)        # rest of the episode
    assert sh.step_count == int(fx["steps"])
    assert np.array_equal(np.asarray(sh.agents_time), fx["final_atime"])
    assert [float(success), total_time, ttime] == list(fx["result"][:3])


def test_alan_window_period_matches_the_reference_run():
    """all weights drop to zero on steps 121, 242, ... (except the action just chosen)."""
    from collision_avoidance_b200.alan import alan_window_steps
    fx = load_alan("shell_alan_circle32")
    period = alan_window_steps(1 / 60., 2)
    assert period == 121
    w = fx["w"]
    for t in range(1, w.shape[0]):
        nonzero_per_agent = (w[t] != 0).sum(-1)
        if t % period == 0:
            assert (nonzero_per_agent <= 1).all(), t
    assert ((w[period - 1] != 0).sum(-1) > 1).any()


# ------------------------------------------------------------------------------ gym env shell (a10-a13, a15)
@pytest.mark.parametrize("name", ["shell_env", "shell_env_small"])
def test_env_shell_oracle_equals_reference_env(name):
    """collision_avoidence_env.py:367-416 (step), :461-488 (reset), :447-458 (orca_step),
    :231-318 (_get_obs): observations, rewards, dones, states -- exactly."""
    fx = load_env(name)
    sh = oracle_from_env_fixture(fx)
    kinds = fx["kind"]
    for r in range(1, len(kinds)):
        if kinds[r] == 1:
            obs, rew, done = sh.step(fx["theta"][r])
            assert np.array_equal(np.asarray(rew), fx["rew"][r]), r
            assert bool(done) == bool(fx["done"][r]), r
        elif kinds[r] == 2:
            obs = sh.reset(fx["pos"][r])
        else:
            obs = sh.orca_step()
        assert np.array_equal(np.asarray(obs), fx["obs"][r]), r
        assert np.array_equal(sh.sim.positions(), fx["pos"][r]), r
        assert np.array_equal(sh.sim.velocities(), fx["vel"][r]), r
        assert sh.agents_done == list(fx["agents_done"][r]), r
    assert fx["agents_done"].max() == 1, "the recorded episode must exercise the done / target switch"
    if name == "shell_env_small":
        assert fx["done"].max() == 1, "the small world must run to done['__all__']"


# ------------------------------------------------------------------------------ rvo2 boundary (b)
@pytest.mark.parametrize("trace", ["alan_blocks_online", "alan_congested_orca", "env_step_reset_orca"])
def test_boundary_call_trace_replays_on_the_oracle(trace):
    calls = load_calltrace()[trace]
    n, worst = replay_calltrace(calls, rvo2_oracle.PyRVOSimulator, atol=0.0)
    assert n > 1000 and worst == 0.0


# ------------------------------------------------------------------------------ MCMC trainer (f1)
def _mcmc_fixture():
    with open(os.path.join(GOLDEN, "shell_mcmc.json")) as f:
        return json.load(f)


def _fake_cost(actions):
    ang = np.arctan2([a[1] for a in actions], [a[0] for a in actions])
    return float(10.0 + np.sum(np.cos(3.0 * ang)) + 0.25 * len(actions))


def test_mcmc_moves_equal_the_reference_moves():
    """Train_ALAN_action_space.py:70-126,133-135 under the reference's random streams."""
    from collision_avoidance_b200.mcmc import MCMC_trainer
    fx = _mcmc_fixture()
    tr = MCMC_trainer(numRounds=2, seed=501, cost_fn=_fake_cost, reference_semantics=True)
    actions = [list(map(float, a)) for a in fx["moves"][0]["before"]]
    assert [tuple(a) for a in actions] == [tuple(map(float, a)) for a in tr.actions[0]]   # (1,0) + random_action()
    actions = [tuple(a) for a in actions]
    kinds = set()
    for k, mv in enumerate(fx["moves"]):
        modification = tr.select_modification(actions, k)
        if modification == 1 and len(actions) <= 2:
            modification = 0
        assert modification == mv["modification"], k
        d, actions = tr.apply_modification(actions, modification)
        assert d == mv["dist"], k
        assert [list(a) for a in actions] == mv["after"], k
        assert tr.symmetric_likelihood(d) == pytest.approx(mv["likelihood"], rel=1e-15)
        kinds.add(modification)
    assert kinds == {0, 1, 2}


@pytest.mark.parametrize("idx", [0, 1])
def test_mcmc_train_equals_the_reference_train(idx):
    """Train_ALAN_action_space.py:27-47: proposals, evaluations, the accept rule, the (rising)
    temperature and the aliasing of the working / best sets, on a deterministic cost."""
    from collision_avoidance_b200.mcmc import MCMC_trainer
    fx = _mcmc_fixture()["train"][idx]
    tr = MCMC_trainer(numRounds=fx["numRounds"], seed=fx["seed"], cost_fn=_fake_cost, reference_semantics=True)
    seen = []
    cost = tr.cost_fn
    tr.cost_fn = lambda a: (seen.append([list(x) for x in a]), cost(a))[1]
    best = tr.train()
    assert seen == [h["proposal"] for h in fx["history"]]
    assert [list(a) for a in best] == fx["actions_opt"]
    assert tr.eval_opt[0] == fx["eval_opt"] and tr.eval[0] == fx["final_eval"]
    assert [list(a) for a in tr.actions[0]] == fx["final_actions"]
    assert tr.temp == pytest.approx(fx["final_temp"], rel=1e-12) and tr.temp > 0.9   # the reference heats up


def test_mcmc_default_semantics_fix_the_reference_bugs():
    from collision_avoidance_b200.mcmc import MCMC_trainer
    tr = MCMC_trainer(numRounds=30, seed=7, cost_fn=_fake_cost, chains=3)
    start = [list(a) for a in tr.actions]
    best = tr.train()
    assert tr.temp < 0.9 + 1e-9 and tr.temp == pytest.approx(0.1 + tr.delta_temp, abs=1e-9)
    assert min(tr.eval_opt) <= min(_fake_cost(a) for a in start)
    assert _fake_cost(best) == pytest.approx(min(tr.eval_opt))
    assert all(a is not b for a, b in zip(tr.actions, tr.actions_opt))
