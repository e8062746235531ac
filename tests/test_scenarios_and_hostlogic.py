"""Host-side logic: scenario geometry vs. the reference's formulas, .act round trip,
ALAN reset period."""
import json
import os
from math import pi, sqrt

import numpy as np
import pytest

from collision_avoidance_b200 import actfile, scenarios
from collision_avoidance_b200.alan import alan_window_steps, unit_actions

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_circle_geometry_matches_reference_formulas():
    # ALAN_true.py:299-301: R = 0.5*3*N/(2 pi); envsize = 2R + 4r (N=16 -> 3.8197, 9.6394)
    scn = scenarios.circle(3, 16, seed=0, rotate=False)
    R = 1.5 * 16 / (2 * pi)
    assert scn.envsize == pytest.approx(2 * R + 2)
    assert R == pytest.approx(3.8197, abs=1e-4) and scn.envsize == pytest.approx(9.6394, abs=1e-4)
    c = scn.envsize / 2
    assert np.allclose(np.linalg.norm(scn.pos - c, axis=-1), R, atol=1e-5)
    assert np.allclose(scn.pos + scn.goal, 2 * c, atol=1e-5)            # antipodal goals
    assert np.allclose(np.linalg.norm(scn.vel, axis=-1), 1.0, atol=1e-6)  # Q2: unit initial velocity
    assert scn.obstacles == [[(0.0, 0.0), (0.0, scn.envsize), (scn.envsize, scn.envsize), (scn.envsize, 0.0)]]
    assert scenarios.circle(1, 32).envsize == pytest.approx(17.2789, abs=1e-4)


def test_crowd_and_default_env_geometry():
    scn = scenarios.crowd(2, 256, seed=1, blocks=4)
    assert scn.envsize == pytest.approx(32.0) and scn.per_env_obstacles and len(scn.obstacles[0]) == 5
    assert scenarios.crowd(1, 1_000_0, seed=1).envsize == pytest.approx(2 * sqrt(1e4))
    d = scenarios.default_env(4, 10, seed=2)
    assert d.params["maxNeighbors"] == 5 and d.params["neighborDist"] == 1.5
    assert (d.pos[..., 0] >= 5).all() and (d.pos[..., 0] <= 10).all()
    assert (d.goal == np.array([1.0, 5.0], np.float32)).all() and (d.goal2 == np.array([-10.0, 5.0], np.float32)).all()
    assert len(d.obstacles) == 3 and d.obstacles[0][0] == (-15.0, 0.0)
    with pytest.raises(ValueError):
        scenarios.make("nope", 1, 4)


def test_all_scenarios_build():
    for name in ("circle", "crowd", "blocks", "congested", "incoming", "deadlock", "default_env"):
        s = scenarios.make(name, 2, 17 if name == "incoming" else 12, seed=3)
        assert s.pos.shape == (2, s.agents_per_env, 2) and s.pos.dtype == np.float32
        assert s.goal2.shape == s.goal.shape


def test_act_files_round_trip():
    with open(os.path.join(GOLD, "act_tables.json")) as f:
        tables = json.load(f)
    for name, t in tables.items():
        acts = [tuple(a) for a in t]
        text = actfile.dumps(acts)
        assert actfile.loads(text) == acts
    assert actfile.dumps([(1, 0), (0.5, -0.5)]) == "[(1, 0), (0.5, -0.5)]"   # str(list) like the reference
    with pytest.raises(ValueError):
        actfile.loads("[]")


def test_alan_constants():
    assert alan_window_steps(1 / 60., 2) == 121
    assert alan_window_steps(0.25, 2) == 8
    u = unit_actions([(1, 0), (0.70711, 0.70711), (0, -3)])
    assert np.allclose(u, [[1, 0], [sqrt(0.5), sqrt(0.5)], [0, -1]], atol=1e-6)
