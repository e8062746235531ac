#!/usr/bin/env python
"""Policy network forward pass (row f2) on 1M observation rows: time, FP32 FMA rate, HBM GB/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from collision_avoidance_b200.policy import SharedMLPPolicy  # noqa: E402
from collision_avoidance_b200.sim import BatchedRVOSimulator  # noqa: E402

rows = 1_000_000
sim = BatchedRVOSimulator(2, 8, 1 / 60., 5.0, 10, 1.5, 1.5, 0.5, 1.0)
obs = torch.randn(rows, 64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for impl in ("tcgen05", "fp32"):
    pol = SharedMLPPolicy(sim, impl=impl)
    for _ in range(3):
        pol(obs)
    ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pol(obs)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    fma = rows * (64 * 64 * 2 + 64 * 2)
    print(json.dumps({"kernel": "policy_mlp_tc_kernel" if impl == "tcgen05" else "policy_mlp_kernel", "rows": rows, "ms": ms,
                      "useful_tfma_per_s": fma / ms / 1e9, "fp32_peak_tfma_per_s": 148 * 128 * 1.965e9 / 1e12,
                      "hbm_gbs_algorithmic": rows * (256 + 8) / ms / 1e6, "rows_per_s": rows / ms * 1e3}))
