"""Shared helpers for the parity tests: build oracle simulators from a Scenario and compare."""
from __future__ import annotations

import numpy as np

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
