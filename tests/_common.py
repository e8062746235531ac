"""Shared helpers for the parity tests: build oracle simulators from a Scenario and compare."""
from __future__ import annotations

import numpy as np

from collision_avoidance_b200 import scenarios

from oracle.helpers import goal_pref, oracle_sims  # noqa: F401
from oracle.rvo2_oracle import PyRVOSimulator as OraclePyRVO  # noqa: F401


def camel(p):
    return dict(timeStep=p["time_step"], neighborDist=p["neighbor_dist"], maxNeighbors=p["max_neighbors"],
                timeHorizon=p["time_horizon"], timeHorizonObst=p["time_horizon_obst"], radius=p["radius"],
                maxSpeed=p["max_speed"])


def snake(p):
    return dict(time_step=p["timeStep"], neighbor_dist=p["neighborDist"], max_neighbors=p["maxNeighbors"],
                time_horizon=p["timeHorizon"], time_horizon_obst=p["timeHorizonObst"], radius=p["radius"],
                max_speed=p["maxSpeed"])


def neighbor_sets_equal_up_to_ties(ids_a, ids_b, dsq_of):
    """Neighbor lists agree exactly, or differ only among candidates at bit-equal distance."""
    if list(ids_a) == list(ids_b):
        return True
    if len(ids_a) != len(ids_b):
        return False
    da = sorted(dsq_of(i) for i in ids_a)
    db = sorted(dsq_of(i) for i in ids_b)
    return da == db


def pillar_hall(num_agents=10, seed=0, pillars=9, side=0.16, ring=1.15):
    """A world that overflows the fast path's fixed capacities (16 obstacle neighbors, 6 obstacle
    lines per agent): a ring of small square pillars with some agents inside it (every pillar gives
    them a half-plane of its own direction) and some outside.  9 pillars = 36 + 4 vertices (+ BSP
    splits) <= ORCA_SLOW_MAX_OBST, so no agent can exceed the slow path's capacity either."""
    rng = np.random.default_rng(seed)
    polys = [[(-6.0, -6.0), (-6.0, 6.0), (6.0, 6.0), (6.0, -6.0)]]            # clockwise wall
    h = side / 2
    for k in range(pillars):
        cx, cy = ring * np.cos(2 * np.pi * k / pillars + 0.1), ring * np.sin(2 * np.pi * k / pillars + 0.1)
        polys.append([(cx - h, cy - h), (cx + h, cy - h), (cx + h, cy + h), (cx - h, cy + h)])   # counter-clockwise block
    inside = num_agents // 2
    ang = rng.uniform(0, 2 * np.pi, num_agents)
    r = np.concatenate([rng.uniform(0.0, 0.35, inside), rng.uniform(2.2, 3.0, num_agents - inside)])
    pos = np.stack([r * np.cos(ang), r * np.sin(ang)], -1)[None].astype(np.float32)
    goal = np.stack([3.0 * np.cos(ang + 2.5), 3.0 * np.sin(ang + 2.5)], -1)[None].astype(np.float32)
    goal[0, inside:] = -pos[0, inside:] * 0.1
    vel = np.zeros_like(pos)
    return scenarios.Scenario("pillar_hall", pos, vel, goal, goal.copy(), 12.0, polys)
