#!/usr/bin/env python
"""How the placement of the pinned host buffers moves orca_step_host (dev tool).
usage: tools/e2e_hostmem.py torch|thp|wc    (one process per trial: placement is decided at allocation)"""
import ctypes
import mmap
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from collision_avoidance_b200 import _lib, scenarios  # noqa: E402
from collision_avoidance_b200.sim import BatchedRVOSimulator  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "torch"
E, N = 65536, 16
scn = scenarios.circle(E, N, seed=1234)
sim = BatchedRVOSimulator(E, N, **scn.params)
sim.set_obstacles(scn.obstacles)
cudart = torch.cuda.cudart()
keep = []


def host_buffer(src: np.ndarray, write_combined=False):
    if mode == "torch" or (mode == "wc" and not write_combined):
        return torch.from_numpy(src.copy()).pin_memory()
    nbytes = src.nbytes
    if mode == "thp":
        size = (nbytes + (2 << 20) - 1) & ~((2 << 20) - 1)
        m = mmap.mmap(-1, size + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
        aligned = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        rc = libc.madvise(ctypes.c_void_p(aligned), size, 14)  # MADV_HUGEPAGE
        arr = np.frombuffer(m, dtype=np.float32, count=src.size, offset=aligned - addr).reshape(src.shape)
        arr[...] = src  # touch
        t = torch.from_numpy(arr)
        err = cudart.cudaHostRegister(aligned, size, 1 | 2)  # portable | mapped
        assert int(err) == 0, err
        keep.append(m)
        return t
    if mode == "wc":
        lib = ctypes.CDLL("libcudart.so.12")
        p = ctypes.c_void_p()
        rc = lib.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(1 | 2 | 4))  # portable | mapped | write-combined
        assert rc == 0, rc
        buf = (ctypes.c_float * src.size).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float32).reshape(src.shape)
        arr[...] = src
        return torch.from_numpy(arr)
    raise SystemExit("mode?")


pos_h, vel_h, goal_h = host_buffer(scn.pos), host_buffer(scn.vel), host_buffer(scn.goal, write_combined=True)
sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=True, steps=150)
for _ in range(8):
    sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
ts = []
for _ in range(200):
    t0 = time.perf_counter()
    sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
    ts.append(time.perf_counter() - t0)
ts.sort()
thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists("/sys/kernel/mm/transparent_hugepage/enabled") else "?"
print(f"{mode:6s} median {ts[100] * 1e3:.3f} ms  min {ts[0] * 1e3:.3f}  p90 {ts[180] * 1e3:.3f}   checksum {float(pos_h.sum()):.3f}   thp={thp}", flush=True)
