"""Test-only driver of tests/host_emul/emul.cpp (the library's per-agent device code compiled
for the host).  Lets the CPU suite check kernel LOGIC against the oracle without a GPU."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "host_emul", "emul.cpp")
_LIB = os.path.join(_HERE, "host_emul", "libemul.so")
_CSRC = os.path.join(_HERE, "..", "collision_avoidance_b200", "csrc")

c_f, c_i, c_p = ctypes.c_float, ctypes.c_int, ctypes.c_void_p


class StepArgs(ctypes.Structure):
    """Mirror of orca::StepArgs (csrc/orca_step_small.cuh)."""
    _fields_ = [
        ("E", c_i), ("N", c_i), ("envs_per_block", c_i), ("env_base", c_i), ("k", c_i),
        ("dt", c_f), ("inv_dt", c_f), ("nd_sq", c_f), ("inv_th", c_f), ("inv_tho", c_f), ("radius", c_f),
        ("vmax", c_f), ("obst_range_sq", c_f),
        ("pos", c_p), ("vel", c_p), ("pref", c_p), ("goal", c_p), ("goal2", c_p),
        ("action_theta", c_p), ("rl_scale", c_f), ("done_x", c_f),
        ("alan_w", c_p), ("alan_actions", c_p), ("alan_action_out", c_p), ("alan_uniform_in", c_p),
        ("A", c_i), ("alan_window", c_i), ("alan_gamma", c_f), ("alan_inv_temp", c_f),
        ("seed", ctypes.c_ulonglong), ("alan_A_env", c_p), ("alan_env_stride", c_i),
        ("reward", c_p), ("done", c_p), ("arrival", c_p), ("env_step", c_p), ("env_done_cnt", c_p),
        ("done_mode", c_i),
        ("nbr_idx", c_p), ("nbr_dsq", c_p), ("nbr_cnt", c_p), ("onbr_idx", c_p), ("onbr_cnt", c_p),
        ("stats", c_p),
        ("vert_pd", c_p), ("vert_link", c_p), ("bsp", c_p), ("bsp_seg", c_p), ("env_nodes", c_p),
        ("shared_nodes", c_i), ("vert_stride", c_i), ("world_slots", c_i), ("world_verts", c_i), ("neighbors_only", c_i), ("grid_path", c_i),
        ("pos_mirror", c_p), ("vel_mirror", c_p), ("tile_grid_inv_cell", c_f),
        ("nbr_hint", c_p), ("hint_slack", c_f), ("cull_rows", c_p), ("cull_geo", c_p),
    ]


class ObsArgs(ctypes.Structure):
    """Mirror of orca::ObsArgs (csrc/orca_obs.cuh)."""
    _fields_ = [
        ("E", c_i), ("N", c_i), ("k", c_i), ("R", c_i), ("C", c_i),
        ("pos", c_p), ("vel", c_p), ("goal", c_p), ("nbr_idx", c_p), ("nbr_cnt", c_p), ("onbr_idx", c_p),
        ("onbr_cnt", c_p), ("vert_pd", c_p), ("vert_link", c_p), ("vert_stride", c_i), ("obs", c_p),
        ("ray_end", c_f * 64), ("poly", c_f * 32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)]
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(d) for d in deps):
            # ORCA_EMUL_SANITIZE=1: AddressSanitizer + UBSan build of the library's per-agent device code
            # (run the suite with LD_PRELOAD=$(gcc -print-file-name=libasan.so); see profiles/README.md)
            san = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g"] if os.environ.get("ORCA_EMUL_SANITIZE") else []
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", *san, *os.environ.get("ORCA_EMUL_DEFINES", "").split(),
                                   "-Wno-unknown-pragmas", "-I" + os.path.join(_HERE, "host_emul"), "-o", _LIB, _SRC])
        L = ctypes.CDLL(_LIB)
        L.emul_stepargs_size.restype = c_i
        assert L.emul_stepargs_size() == ctypes.sizeof(StepArgs), (L.emul_stepargs_size(), ctypes.sizeof(StepArgs))
        assert L.emul_obsargs_size() == ctypes.sizeof(ObsArgs), (L.emul_obsargs_size(), ctypes.sizeof(ObsArgs))
        L.emul_observe.argtypes = [ctypes.POINTER(ObsArgs), c_f, c_f]
        L.emul_step.argtypes = [ctypes.POINTER(StepArgs), c_i]
        L.emul_step.restype = c_i
        L.emul_step_grid.argtypes = [ctypes.POINTER(StepArgs), c_i]
        L.emul_step_grid.restype = c_i
        L.emul_build_world.restype = c_i
        L.emul_philox_uniform.restype = c_f
        L.emul_philox_uniform.argtypes = [ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


class World:
    """Processed obstacle tables of one shared world (host arrays)."""

    def __init__(self, polygons, obst_range=2.0):
        """obst_range = timeHorizonObst * maxSpeed + radius (2.0 for every world of the reference)."""
        L = lib()
        xy = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32).reshape(-1, 2) for p in polygons])
                                  if polygons else np.zeros((0, 2), np.float32))
        sizes = np.asarray([len(p) for p in polygons], np.int32)
        max_v = 4 * max(1, xy.shape[0])
        self.pd = np.zeros((max_v, 4), np.float32)
        self.link = np.zeros((max_v, 4), np.int32)
        self.bsp = np.zeros((max_v, 4), np.int32)
        self.seg = np.zeros((max_v, 4), np.float32)
        depth = c_i(0)
        self.cull_rows = np.zeros(32, np.uint32)
        self.cull_geo = np.zeros(4, np.float32)
        nv = L.emul_build_world(c_p(xy.ctypes.data), c_p(sizes.ctypes.data), len(polygons), max_v,
                                c_p(self.pd.ctypes.data), c_p(self.link.ctypes.data), c_p(self.bsp.ctypes.data),
                                c_p(self.seg.ctypes.data), ctypes.byref(depth), c_f(obst_range),
                                c_p(self.cull_rows.ctypes.data), c_p(self.cull_geo.ctypes.data))
        assert nv >= 0, nv
        self.nv = nv
        self.depth = depth.value


def emul_step(params, pos, vel, *, policy=0, pref=None, goal=None, goal2=None, world=None, action_theta=None,
              rl_scale=0.3, done_x=2.0, alan_w=None, alan_actions=None, alan_uniform=None, alan_window=121,
              alan_gamma=0.6, alan_temp=0.2, seed=0, done_mode=0, agent_done=None, arrival=None, env_step=None,
              env_done_cnt=None, want_neighbors=False, neighbors_only=False, stats=None, grid=False, tile_grid=None,
              nbr_hint=None):
    """Run one fused step on host arrays, in place.  pos/vel: float32 [E, N, 2].
    Returns dict with optional outputs (reward, action, nbr_idx, nbr_cnt, ...)."""
    E, N = pos.shape[0], pos.shape[1]
    a = StepArgs()
    a.E, a.N, a.envs_per_block, a.k = E, N, 1, int(params["max_neighbors"])
    f32 = np.float32
    # as launch_small_kp (csrc/orca_api.cu): worlds of more than 32 agents search an in-block grid
    if tile_grid is None:
        tile_grid = N > 32 and N <= 256 and not grid
    nd = f32(params["neighbor_dist"])
    a.tile_grid_inv_cell = f32(1.0) / (nd * f32(1.001)) if (tile_grid and nd > 0) else f32(0.0)
    dt = f32(params["time_step"])
    a.dt = dt
    a.inv_dt = f32(1.0) / dt
    a.nd_sq = f32(params["neighbor_dist"]) * f32(params["neighbor_dist"])
    a.inv_th = f32(1.0) / f32(params["time_horizon"])
    a.inv_tho = f32(1.0) / f32(params["time_horizon_obst"])
    a.radius = f32(params["radius"])
    a.vmax = f32(params["max_speed"])
    orange = f32(params["time_horizon_obst"]) * f32(params["max_speed"]) + f32(params["radius"])
    a.obst_range_sq = orange * orange
    out = {}
    keep = [pos, vel, pref, goal, goal2, action_theta, alan_w, alan_actions, alan_uniform, agent_done, arrival,
            env_step, env_done_cnt, stats]
    for arr in keep:
        if arr is not None:
            assert arr.flags["C_CONTIGUOUS"]
    a.pos, a.vel, a.pref, a.goal, a.goal2 = _ptr(pos), _ptr(vel), _ptr(pref), _ptr(goal), _ptr(goal2)
    a.action_theta, a.rl_scale, a.done_x = _ptr(action_theta), rl_scale, done_x
    a.alan_w, a.alan_actions, a.alan_uniform_in = _ptr(alan_w), _ptr(alan_actions), _ptr(alan_uniform)
    if alan_actions is not None:
        a.A = alan_actions.shape[0]
        out["action"] = np.zeros((E, N), np.uint8)
        a.alan_action_out = _ptr(out["action"])
    a.alan_window, a.alan_gamma, a.alan_inv_temp, a.seed = alan_window, alan_gamma, f32(1.0) / f32(alan_temp), seed
    if policy in (2, 3):
        out["reward"] = np.zeros((E, N), np.float32)
        a.reward = _ptr(out["reward"])
    a.done, a.arrival, a.env_step, a.env_done_cnt = _ptr(agent_done), _ptr(arrival), _ptr(env_step), _ptr(env_done_cnt)
    a.done_mode = done_mode
    if want_neighbors or neighbors_only:
        out["nbr_idx"] = np.full((E, N, a.k), -1, np.int32)
        out["nbr_dsq"] = np.zeros((E, N, a.k), np.float32)
        out["nbr_cnt"] = np.zeros((E, N), np.int32)
        out["onbr_idx"] = np.full((E, N, 16), -1, np.int32)
        out["onbr_cnt"] = np.zeros((E, N), np.int32)
        a.nbr_idx, a.nbr_dsq, a.nbr_cnt = _ptr(out["nbr_idx"]), _ptr(out["nbr_dsq"]), _ptr(out["nbr_cnt"])
        a.onbr_idx, a.onbr_cnt = _ptr(out["onbr_idx"]), _ptr(out["onbr_cnt"])
    a.stats = _ptr(stats)
    if nbr_hint is not None:      # float32 [E, N], persistent between steps: starting thresholds of the search
        assert nbr_hint.dtype == np.float32 and nbr_hint.flags["C_CONTIGUOUS"]
        a.nbr_hint = _ptr(nbr_hint)
        a.hint_slack = f32(2.5) * f32(params["max_speed"]) * dt + f32(1e-4)
    if world is not None and world.nv > 0:
        a.vert_pd, a.vert_link, a.bsp, a.bsp_seg = _ptr(world.pd), _ptr(world.link), _ptr(world.bsp), _ptr(world.seg)
        a.shared_nodes, a.vert_stride = world.nv, 0
        a.cull_rows, a.cull_geo = _ptr(world.cull_rows), _ptr(world.cull_geo)
    a.neighbors_only = 1 if neighbors_only else 0
    rc = lib().emul_step_grid(ctypes.byref(a), policy) if grid else lib().emul_step(ctypes.byref(a), policy)
    assert rc == 0
    return out


def emul_observe(params, pos, vel, goal, nbr, world=None, laser_num=16, circle_approx_num=8):
    """Laser scan of every agent (host emulation of observe_kernel).  nbr: dict from emul_step."""
    E, N = pos.shape[0], pos.shape[1]
    a = ObsArgs()
    a.E, a.N, a.k, a.R, a.C = E, N, max(1, int(params["max_neighbors"])), laser_num, circle_approx_num
    obs = np.zeros((E, N, laser_num * 4), np.float32)
    a.pos, a.vel, a.goal = _ptr(pos), _ptr(vel), _ptr(goal)
    a.nbr_idx, a.nbr_cnt, a.onbr_idx, a.onbr_cnt = (_ptr(nbr["nbr_idx"]), _ptr(nbr["nbr_cnt"]), _ptr(nbr["onbr_idx"]),
                                                    _ptr(nbr["onbr_cnt"]))
    if world is not None and world.nv > 0:
        a.vert_pd, a.vert_link = _ptr(world.pd), _ptr(world.link)
    a.obs = _ptr(obs)
    assert lib().emul_observe(ctypes.byref(a), np.float32(params["neighbor_dist"]), np.float32(params["radius"])) == 0
    return obs
