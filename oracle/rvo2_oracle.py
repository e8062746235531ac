"""ctypes front-end of the CPU oracle (oracle/rvo2_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  Nothing under
``collision_avoidance_b200/`` may import this module.

``PyRVOSimulator`` mirrors the call surface of ``rvo2.PyRVOSimulator`` that the
reference uses (call sites
collision_avoidance/envs/collision_avoidence_env.py:62-68,126-148,385 and
collision_avoidance/ALAN/ALAN_true.py:22-28,461-479,601): tuples in, tuples out,
sequential ids.  PARITY UNPINNED vs. upstream rvo2 (module not available);
pinned vs. analytic known answers (tests/test_oracle_known_answers.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librvo2_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ -O2 -ffp-contract=off)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "rvo2_oracle.cpp"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = ctypes.CDLL(_LIB_PATH)
    f, i, p = ctypes.c_float, ctypes.c_int, ctypes.c_void_p
    fp = ctypes.POINTER(ctypes.c_float)
    sig = {
        "rvo_create": (p, [f, f, i, f, f, f, f, f, f]),
        "rvo_destroy": (None, [p]),
        "rvo_add_agent": (i, [p, f, f, f, i, f, f, f, f, f, f]),
        "rvo_add_agent_default": (i, [p, f, f]),
        "rvo_add_agents": (i, [p, fp, fp, i]),
        "rvo_add_obstacle": (i, [p, fp, i]),
        "rvo_process_obstacles": (None, [p]),
        "rvo_do_step": (None, [p]),
        "rvo_num_agents": (i, [p]),
        "rvo_num_obstacle_vertices": (i, [p]),
        "rvo_global_time": (f, [p]),
        "rvo_set_agent_pref_velocity": (None, [p, i, f, f]),
        "rvo_set_agent_position": (None, [p, i, f, f]),
        "rvo_set_agent_velocity": (None, [p, i, f, f]),
        "rvo_get_agent_position": (None, [p, i, fp]),
        "rvo_get_agent_velocity": (None, [p, i, fp]),
        "rvo_get_agent_pref_velocity": (None, [p, i, fp]),
        "rvo_get_agent_num_agent_neighbors": (i, [p, i]),
        "rvo_get_agent_agent_neighbor": (i, [p, i, i]),
        "rvo_get_agent_agent_neighbor_distsq": (f, [p, i, i]),
        "rvo_get_agent_num_obstacle_neighbors": (i, [p, i]),
        "rvo_get_agent_obstacle_neighbor": (i, [p, i, i]),
        "rvo_get_agent_obstacle_neighbor_distsq": (f, [p, i, i]),
        "rvo_get_next_obstacle_vertex_no": (i, [p, i]),
        "rvo_get_prev_obstacle_vertex_no": (i, [p, i]),
        "rvo_get_obstacle_vertex": (None, [p, i, fp]),
        "rvo_get_obstacle_vertex_convex": (i, [p, i]),
        "rvo_get_obstacle_vertex_unit_dir": (None, [p, i, fp]),
        "rvo_get_agent_num_orca_lines": (i, [p, i]),
        "rvo_get_agent_num_obst_orca_lines": (i, [p, i]),
        "rvo_get_agent_orca_line": (None, [p, i, i, fp]),
        "rvo_get_positions": (None, [p, fp]),
        "rvo_get_velocities": (None, [p, fp]),
        "rvo_set_positions": (None, [p, fp]),
        "rvo_set_velocities": (None, [p, fp]),
        "rvo_set_pref_velocities": (None, [p, fp]),
        "rvo_solve_lp": (i, [fp, i, i, f, f, f, fp]),
        "rvo_batch_orca_steps": (None, [ctypes.POINTER(p), i, ctypes.POINTER(ctypes.c_double), i, i, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class PyRVOSimulator:
    """Oracle stand-in for ``rvo2.PyRVOSimulator``."""

    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed,
                 velocity=(0.0, 0.0)):
        self._L = lib()
        self._h = self._L.rvo_create(timeStep, neighborDist, int(maxNeighbors), timeHorizon, timeHorizonObst,
                                     radius, maxSpeed, float(velocity[0]), float(velocity[1]))
        self._buf = (ctypes.c_float * 4)()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.rvo_destroy(h)

    @property
    def handle(self):
        return self._h

    # -- construction ------------------------------------------------------
    def addAgent(self, pos, neighborDist=None, maxNeighbors=None, timeHorizon=None, timeHorizonObst=None,
                 radius=None, maxSpeed=None, velocity=None):
        opt = (neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed, velocity)
        if all(o is None for o in opt):
            return self._L.rvo_add_agent_default(self._h, float(pos[0]), float(pos[1]))
        if any(o is None for o in opt):
            raise ValueError("Either pass only 'pos', or pass all parameters.")
        return self._L.rvo_add_agent(self._h, float(pos[0]), float(pos[1]), float(neighborDist), int(maxNeighbors),
                                     float(timeHorizon), float(timeHorizonObst), float(radius), float(maxSpeed),
                                     float(velocity[0]), float(velocity[1]))

    def addObstacle(self, vertices):
        arr = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 2))
        r = self._L.rvo_add_obstacle(self._h, _fptr(arr), int(arr.shape[0]))
        if r < 0:
            raise RuntimeError("Error adding obstacle to RVO simulation")
        return r

    def processObstacles(self):
        self._L.rvo_process_obstacles(self._h)

    def doStep(self):
        self._L.rvo_do_step(self._h)

    # -- scalar getters / setters -----------------------------------------
    def _get2(self, fn, i):
        fn(self._h, int(i), self._buf)
        return (float(self._buf[0]), float(self._buf[1]))

    def getAgentPosition(self, i):
        return self._get2(self._L.rvo_get_agent_position, i)

    def getAgentVelocity(self, i):
        return self._get2(self._L.rvo_get_agent_velocity, i)

    def getAgentPrefVelocity(self, i):
        return self._get2(self._L.rvo_get_agent_pref_velocity, i)

    def setAgentPrefVelocity(self, i, v):
        self._L.rvo_set_agent_pref_velocity(self._h, int(i), float(v[0]), float(v[1]))

    def setAgentPosition(self, i, p):
        self._L.rvo_set_agent_position(self._h, int(i), float(p[0]), float(p[1]))

    def setAgentVelocity(self, i, v):
        self._L.rvo_set_agent_velocity(self._h, int(i), float(v[0]), float(v[1]))

    def getAgentNumAgentNeighbors(self, i):
        return self._L.rvo_get_agent_num_agent_neighbors(self._h, int(i))

    def getAgentAgentNeighbor(self, i, j):
        return self._L.rvo_get_agent_agent_neighbor(self._h, int(i), int(j))

    def getAgentNumObstacleNeighbors(self, i):
        return self._L.rvo_get_agent_num_obstacle_neighbors(self._h, int(i))

    def getAgentObstacleNeighbor(self, i, j):
        return self._L.rvo_get_agent_obstacle_neighbor(self._h, int(i), int(j))

    def getNextObstacleVertexNo(self, v):
        return self._L.rvo_get_next_obstacle_vertex_no(self._h, int(v))

    def getPrevObstacleVertexNo(self, v):
        return self._L.rvo_get_prev_obstacle_vertex_no(self._h, int(v))

    def getObstacleVertex(self, v):
        return self._get2(self._L.rvo_get_obstacle_vertex, v)

    def getNumAgents(self):
        return self._L.rvo_num_agents(self._h)

    def getNumObstacleVertices(self):
        return self._L.rvo_num_obstacle_vertices(self._h)

    def getGlobalTime(self):
        return float(self._L.rvo_global_time(self._h))

    # -- oracle-only bulk helpers (float32 arrays [n,2]) --------------------
    def add_agents(self, pos, vel):
        """addAgent(pos) with the constructor's defaults for every row; ``vel`` = initial velocities."""
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        return self._L.rvo_add_agents(self._h, _fptr(pos), _fptr(vel), int(pos.shape[0]))

    def positions(self):
        out = np.empty((self.getNumAgents(), 2), np.float32)
        self._L.rvo_get_positions(self._h, _fptr(out))
        return out

    def velocities(self):
        out = np.empty((self.getNumAgents(), 2), np.float32)
        self._L.rvo_get_velocities(self._h, _fptr(out))
        return out

    def set_positions(self, a):
        a = np.ascontiguousarray(a, np.float32)
        self._L.rvo_set_positions(self._h, _fptr(a))

    def set_velocities(self, a):
        a = np.ascontiguousarray(a, np.float32)
        self._L.rvo_set_velocities(self._h, _fptr(a))

    def set_pref_velocities(self, a):
        a = np.ascontiguousarray(a, np.float32)
        self._L.rvo_set_pref_velocities(self._h, _fptr(a))

    def agent_neighbors(self, i):
        """[(agent id, distSq)] of the last doStep, ascending distance."""
        n = self.getAgentNumAgentNeighbors(i)
        return [(self._L.rvo_get_agent_agent_neighbor(self._h, i, j),
                 float(self._L.rvo_get_agent_agent_neighbor_distsq(self._h, i, j))) for j in range(n)]

    def obstacle_neighbors(self, i):
        n = self.getAgentNumObstacleNeighbors(i)
        return [(self._L.rvo_get_agent_obstacle_neighbor(self._h, i, j),
                 float(self._L.rvo_get_agent_obstacle_neighbor_distsq(self._h, i, j))) for j in range(n)]

    def orca_lines(self, i):
        """(lines[n,4] = point.xy, dir.xy ; number of obstacle lines) of the last doStep."""
        n = self._L.rvo_get_agent_num_orca_lines(self._h, int(i))
        out = np.zeros((n, 4), np.float32)
        for j in range(n):
            self._L.rvo_get_agent_orca_line(self._h, int(i), j, self._buf)
            out[j] = [self._buf[k] for k in range(4)]
        return out, self._L.rvo_get_agent_num_obst_orca_lines(self._h, int(i))

    def obstacle_vertex_table(self):
        """Post-processObstacles vertex table: points, unit_dirs, next, prev, convex."""
        n = self.getNumObstacleVertices()
        pts = np.zeros((n, 2), np.float32)
        dirs = np.zeros((n, 2), np.float32)
        nxt = np.zeros(n, np.int32)
        prv = np.zeros(n, np.int32)
        cvx = np.zeros(n, np.int32)
        for v in range(n):
            pts[v] = self.getObstacleVertex(v)
            self._L.rvo_get_obstacle_vertex_unit_dir(self._h, v, self._buf)
            dirs[v] = (self._buf[0], self._buf[1])
            nxt[v] = self.getNextObstacleVertexNo(v)
            prv[v] = self.getPrevObstacleVertexNo(v)
            cvx[v] = self._L.rvo_get_obstacle_vertex_convex(self._h, v)
        return pts, dirs, nxt, prv, cvx


def solve_lp(lines, n_obst, radius, pref):
    """LP2 -> LP3 exactly as computeNewVelocity runs them. Returns (fail_index, result[2])."""
    lines = np.ascontiguousarray(lines, np.float32).reshape(-1, 4)
    out = np.zeros(2, np.float32)
    fail = lib().rvo_solve_lp(_fptr(lines), int(lines.shape[0]), int(n_obst), float(radius), float(pref[0]),
                              float(pref[1]), _fptr(out))
    return fail, out


def batch_orca_steps(sims, goals, steps, threads):
    """CPU-baseline driver: step every simulator ``steps`` times (orca_step policy)."""
    n = len(sims)
    arr = (ctypes.c_void_p * n)(*[s.handle for s in sims])
    goals = np.ascontiguousarray(goals, np.float64)
    agents = goals.shape[1]
    lib().rvo_batch_orca_steps(arr, n, goals.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(agents),
                               int(steps), int(threads))
