"""Loaders for the fixtures recorded from the unmodified reference shells
(tests/golden/make_shell_golden.py) + builders that put a checker or a product object into the
recorded state.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import glob
import gzip
import json
import os

import numpy as np

from collision_avoidance_b200 import scenarios

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def alan_fixture_names(mode):
    pat = "shell_alan_*.npz" if mode == 1 else "shell_orca_*.npz"
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, pat)))


def load_alan(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def load_env(name="shell_env"):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def load_calltrace():
    with gzip.open(os.path.join(GOLDEN, "shell_calltrace.json.gz"), "rt") as f:
        return json.load(f)


def scenario_from_fixture(fx, frame=0, params=None):
    """One-world Scenario holding frame ``frame`` of an ALAN / ORCA fixture."""
    polys = [[tuple(map(float, v)) for v in poly] for poly in fx["polygons"]]
    goal = fx["tgt"][frame].astype(np.float32)[None]
    goal2 = fx["tgt2"].astype(np.float32)[None]
    return scenarios.Scenario(str(fx["scenario"]), fx["pos"][frame][None].copy(), fx["vel"][frame][None].copy(), goal, goal2,
                              float(fx["envsize"]), polys, params=dict(params or scenarios.ALAN_PARAMS))


def oracle_from_alan_fixture(fx, frame=0, sim_cls=None):
    """AlanShellOracle in the state of frame ``frame`` (float64 targets as the reference holds them)."""
    from oracle.shell_oracle import AlanShellOracle
    scn = scenario_from_fixture(fx, frame)
    sh = AlanShellOracle(scn, 0, online_actions=[tuple(a) for a in fx["actions"]], sim_cls=sim_cls)
    sh.targets = [(tuple(fx["tgt"][frame][i]), tuple(fx["tgt2"][i])) for i in range(sh.N)]
    sh.sim.set_pref_velocities(fx["pref"][frame])
    sh.action_weights = [list(w) for w in fx["w"][frame]]
    sh.agents_done = [int(d) for d in fx["done"][frame]]
    sh.agents_time = [float(t) for t in fx["atime"][frame]]
    sh.step_count = frame
    # all action timers advance together (ALAN_true.py:621-625): rebuild them by replaying the additions
    t = 0.0
    for _ in range(frame):
        t += sh.timeStep
        if t >= sh.timewindow:
            t = 0
    sh.action_times = [[t] * len(sh.online_actions) for _ in range(sh.N)]
    return sh


def oracle_from_env_fixture(fx, sim_cls=None):
    """EnvShellOracle right after the reference constructor (which ends with reset())."""
    from oracle.shell_oracle import EnvShellOracle
    n = int(fx["numAgents"])
    streams = scenarios.reference_streams(int(fx["seed"]), 1)
    scn = scenarios.default_env(1, n, reference_rng=streams)
    sh = EnvShellOracle(scn, 0, max_step=int(fx["max_step"]), sim_cls=sim_cls)
    sh.reset(fx["pos"][0])
    return sh


def _close(a, b, atol):
    if isinstance(b, (list, tuple)):
        return max([_close(x, y, atol) for x, y in zip(a, b)] + [0.0]) if len(a) == len(b) else float("inf")
    if isinstance(b, int) and not isinstance(b, bool):
        return 0.0 if int(a) == b else float("inf")
    return abs(float(a) - float(b))


def replay_calltrace(calls, sim_cls, atol=0.0, **ctor_kw):
    """Issue the recorded boundary calls against ``sim_cls``; every recorded return value must be
    reproduced (ids exactly, floats within ``atol``).  Returns (calls issued, worst float error)."""
    sim, worst, n = None, 0.0, 0
    for name, args, kwargs, ret in calls:
        n += 1
        if name == "__init__":
            sim = sim_cls(*args, **kwargs, **ctor_kw)
            continue
        got = getattr(sim, name)(*args, **kwargs)
        if ret is None:
            continue
        err = _close(got, ret, atol)
        if err > atol:
            raise AssertionError(f"call #{n} {name}{tuple(args)}: got {got}, reference shell saw {ret} (err {err})")
        worst = max(worst, err)
    return n, worst
