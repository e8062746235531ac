"""Oracle-side helpers (TEST INFRASTRUCTURE): build one oracle simulator per env of a
scenario the way the reference shells do, and the float64 goal-directed preferred velocity."""
from __future__ import annotations

import numpy as np

from .rvo2_oracle import PyRVOSimulator as OraclePyRVO


def oracle_sims(scn, envs=None):
    """One oracle PyRVOSimulator per env, agents + obstacles added like ALAN_true.py:461-479."""
    P = scn.params
    sims = []
    envs = range(scn.num_envs) if envs is None else envs
    for e in envs:
        s = OraclePyRVO(P["timeStep"], P["neighborDist"], P["maxNeighbors"], P["timeHorizon"], P["timeHorizonObst"],
                        P["radius"], P["maxSpeed"])
        for i in range(scn.agents_per_env):
            s.addAgent(tuple(map(float, scn.pos[e, i])), P["neighborDist"], P["maxNeighbors"], P["timeHorizon"],
                       P["timeHorizonObst"], P["radius"], P["maxSpeed"], tuple(map(float, scn.vel[e, i])))
        polys = scn.obstacles[e] if scn.per_env_obstacles else scn.obstacles
        for poly in polys:
            s.addObstacle([tuple(map(float, v)) for v in poly])
        s.processObstacles()
        sims.append(s)
    return sims


def goal_pref(pos, goal):
    """(cos, sin) of atan2(goal - pos) in float64 (collision_avoidence_env.py:156-162)."""
    d = goal.astype(np.float64) - pos.astype(np.float64)
    ang = np.arctan2(d[..., 1], d[..., 0])
    return np.stack([np.cos(ang), np.sin(ang)], -1)
