"""Scalar drop-in for ``rvo2.PyRVOSimulator`` (one world) on top of the batched CUDA simulator.

The reference shells talk to RVO2 through 16 methods with Python tuples in and out
(collision_avoidence_env.py:62-68,123-148,154-157,237-252,283-312,385,479 and
ALAN_true.py:22-28,461-479,490,553,598-613).  This class offers exactly those, so
``sys.modules['rvo2'] = collision_avoidance_b200.rvo2_compat`` lets the reference's shell logic
run unmodified against the CUDA path (tests/test_gpu_compat.py replays the recorded call trace of the unmodified shells against it).  It is a
compatibility shim, not the fast path: every ``doStep`` is one kernel launch for one world plus
host<->device mirrors; throughput comes from ``BatchedRVOSimulator`` / ``envs`` / ``alan``.

Restrictions (checked, never silent): all agents of a simulator share their parameters (the
reference always passes the same ones), ``maxNeighbors <= 16``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .sim import BatchedRVOSimulator


class PyRVOSimulator:
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed,
                 velocity=(0.0, 0.0), device="cuda:0"):
        self._time_step = float(timeStep)
        self._defaults = (float(neighborDist), int(maxNeighbors), float(timeHorizon), float(timeHorizonObst),
                          float(radius), float(maxSpeed))
        self._default_velocity = (float(velocity[0]), float(velocity[1]))
        self._device = device
        self._agent_params: Optional[Tuple] = None
        self._pos: List[Tuple[float, float]] = []
        self._vel: List[Tuple[float, float]] = []
        self._pref: List[Tuple[float, float]] = []
        self._polygons: List[np.ndarray] = []
        self._processed = False
        self._sim: Optional[BatchedRVOSimulator] = None
        self._dirty = True            # host state newer than device state
        self._nbr_host = None         # (nbr_idx, nbr_cnt, obst_idx, obst_cnt) of the last doStep
        self._global_time = 0.0
        self._vertex_table = None

    # ------------------------------------------------------------------ construction
    def addAgent(self, pos, neighborDist=None, maxNeighbors=None, timeHorizon=None, timeHorizonObst=None,
                 radius=None, maxSpeed=None, velocity=None):
        opt = (neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed, velocity)
        if all(o is None for o in opt):
            params, vel = self._defaults, self._default_velocity
        elif any(o is None for o in opt):
            raise ValueError("Either pass only 'pos', or pass all parameters.")
        else:
            params = (float(neighborDist), int(maxNeighbors), float(timeHorizon), float(timeHorizonObst),
                      float(radius), float(maxSpeed))
            vel = (float(velocity[0]), float(velocity[1]))
        if self._agent_params is None:
            self._agent_params = params
        elif params != self._agent_params:
            raise NotImplementedError("all agents of a simulator must share their parameters "
                                      f"(first agent: {self._agent_params}, this one: {params})")
        self._pos.append((float(np.float32(pos[0])), float(np.float32(pos[1]))))
        self._vel.append((float(np.float32(vel[0])), float(np.float32(vel[1]))))
        self._pref.append((0.0, 0.0))
        self._sim = None  # shape changed: rebuild lazily
        self._dirty = True
        return len(self._pos) - 1

    def addObstacle(self, vertices):
        arr = np.asarray(vertices, dtype=np.float32).reshape(-1, 2)
        if arr.shape[0] < 2:
            raise RuntimeError("Error adding obstacle to RVO simulation")
        first = sum(p.shape[0] for p in self._polygons)
        self._polygons.append(arr)
        self._processed = False
        self._vertex_table = None
        return first

    def processObstacles(self):
        self._processed = True
        self._vertex_table = None
        if self._sim is not None:
            self._sim.set_obstacles([p for p in self._polygons])

    # ------------------------------------------------------------------ device plumbing
    def _ensure_sim(self) -> BatchedRVOSimulator:
        if self._sim is None:
            if not self._pos:
                raise RuntimeError("no agents in the simulation")
            nd, k, th, tho, r, vmax = self._agent_params
            self._sim = BatchedRVOSimulator(1, len(self._pos), self._time_step, nd, k, th, tho, r, vmax,
                                            device=self._device)
            if self._processed:
                self._sim.set_obstacles([p for p in self._polygons])
            self._dirty = True
        return self._sim

    def _vertices(self):
        """Vertex table after processObstacles (BSP splits may have appended vertices)."""
        if self._vertex_table is None:
            if self._processed:
                sim = self._ensure_sim() if self._pos else None
                if sim is None:
                    # no agents yet: build a throw-away 1-agent world just to run the BSP
                    tmp = BatchedRVOSimulator(1, 1, self._time_step, *self._defaults, device=self._device)
                    tmp.set_obstacles([p for p in self._polygons])
                    self._vertex_table = tmp.obstacle_vertices(0)
                    tmp.close()
                else:
                    self._vertex_table = sim.obstacle_vertices(0)
            else:
                pts = np.concatenate(self._polygons) if self._polygons else np.zeros((0, 2), np.float32)
                nxt, prv, off = [], [], 0
                for p in self._polygons:
                    n = p.shape[0]
                    nxt += [off + (i + 1) % n for i in range(n)]
                    prv += [off + (i - 1) % n for i in range(n)]
                    off += n
                self._vertex_table = (pts, np.asarray(nxt, np.int32), np.asarray(prv, np.int32),
                                      np.ones(len(nxt), np.int32))
        return self._vertex_table

    def doStep(self):
        sim = self._ensure_sim()
        if self._dirty:
            sim.pos.copy_(torch.tensor(self._pos, dtype=torch.float32).reshape(1, -1, 2))
            sim.vel.copy_(torch.tensor(self._vel, dtype=torch.float32).reshape(1, -1, 2))
        sim.pref.copy_(torch.tensor(self._pref, dtype=torch.float32).reshape(1, -1, 2))
        sim.env_step(policy=_lib.POLICY_EXTERNAL, want_neighbors=True, collect_stats=False)
        pos = sim.pos[0].cpu().numpy()
        vel = sim.vel[0].cpu().numpy()
        self._pos = [(float(p[0]), float(p[1])) for p in pos]
        self._vel = [(float(v[0]), float(v[1])) for v in vel]
        self._nbr_host = (sim.nbr_idx[0].cpu().numpy(), sim.nbr_cnt[0].cpu().numpy(),
                          sim.obst_nbr_idx[0].cpu().numpy(), sim.obst_nbr_cnt[0].cpu().numpy())
        self._dirty = False
        self._global_time += self._time_step

    # ------------------------------------------------------------------ getters / setters
    def getAgentPosition(self, i):
        return self._pos[i]

    def getAgentVelocity(self, i):
        return self._vel[i]

    def getAgentPrefVelocity(self, i):
        return self._pref[i]

    def setAgentPrefVelocity(self, i, v):
        self._pref[i] = (float(np.float32(v[0])), float(np.float32(v[1])))

    def setAgentPosition(self, i, p):
        self._pos[i] = (float(np.float32(p[0])), float(np.float32(p[1])))
        self._dirty = True

    def setAgentVelocity(self, i, v):
        self._vel[i] = (float(np.float32(v[0])), float(np.float32(v[1])))
        self._dirty = True

    def getAgentNumAgentNeighbors(self, i):
        return 0 if self._nbr_host is None else int(self._nbr_host[1][i])

    def getAgentAgentNeighbor(self, i, j):
        if self._nbr_host is None or j >= int(self._nbr_host[1][i]):
            raise IndexError("neighbor index out of range")
        return int(self._nbr_host[0][i, j])

    def getAgentNumObstacleNeighbors(self, i):
        return 0 if self._nbr_host is None else int(self._nbr_host[3][i])

    def getAgentObstacleNeighbor(self, i, j):
        if self._nbr_host is None or j >= int(self._nbr_host[3][i]):
            raise IndexError("neighbor index out of range")
        return int(self._nbr_host[2][i, j])

    def getNextObstacleVertexNo(self, v):
        return int(self._vertices()[1][v])

    def getPrevObstacleVertexNo(self, v):
        return int(self._vertices()[2][v])

    def getObstacleVertex(self, v):
        p = self._vertices()[0][v]
        return (float(p[0]), float(p[1]))

    def getNumAgents(self):
        return len(self._pos)

    def getNumObstacleVertices(self):
        return int(self._vertices()[0].shape[0])

    def getGlobalTime(self):
        return self._global_time

    def getTimeStep(self):
        return self._time_step
