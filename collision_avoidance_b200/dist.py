"""Multi-GPU plumbing: one process per GPU, env instances sharded, statistics all-reduced.

Worlds never interact (every reference world owns a private simulator:
collision_avoidence_env.py:62, ALAN_true.py:22), so a batch is split into contiguous env
ranges, one per rank, and the step needs NO collective.  The only exchange is one packed
all-reduce per reporting interval that makes the episode statistics (TTime,
ALAN_true.py:125-131) and the ALAN action-value aggregates global.  It is a few hundred bytes,
latency-bound, and is issued once per report -- never per step.  Works over NCCL (GPU tensors)
and gloo (CPU tensors; used by the CPU test-suite).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index: int) -> Optional[list]:
    """Pin this process to the CPU cores next to GPU ``device_index`` (NVML's cpu affinity of the
    device = the cores of its NUMA node / PCIe root).  One process drives one GPU; host buffers it
    pins afterwards are first-touched on that node, so the per-step host<->device traffic of
    ``orca_step_host`` does not cross the socket interconnect when several ranks run at once.
    Returns the cores bound to, or None when NVML or the affinity call is unavailable (no-op)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device_index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if device_index < len(ids) and ids[device_index].isdigit():
                    idx = int(ids[device_index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cores = [w * 64 + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1]
        finally:
            pynvml.nvmlShutdown()
        allowed = os.sched_getaffinity(0)
        cores = sorted(c for c in cores if c in allowed)
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


def shard_range(num_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous env range [start, start + count) owned by ``rank``; sizes differ by at most 1."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(int(num_envs), int(world_size))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


@dataclass
class EpisodeSummary:
    agents: float          # agents covered
    finished: float        # agents that reached their goal
    collisions: float
    lp3_calls: float
    agent_steps: float
    mean_time: float       # mean arrival time, unfinished agents counted at max_time
    std_time: float
    ttime: float           # mean + 3 sigma over ALL agents of ALL worlds (ALAN_true.py:125-131, pooled)
    mean_world_ttime: float  # mean over worlds of the per-world TTime (the MCMC cost, Train_ALAN_action_space.py:55-67)
    action_picks: torch.Tensor  # [A] how often each action is currently selected
    action_value: torch.Tensor  # [A] mean last-reward weight per action


def pack_stats(agents_time: torch.Tensor, agents_done: torch.Tensor, max_time: float, sim_stats: torch.Tensor,
               agent_steps: int, action_weights: Optional[torch.Tensor] = None,
               action_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One float64 vector holding every additive statistic of this rank:
    [agents, finished, collisions, lp3, agent_steps, sum t, sum t^2, worlds, sum world TTime,
     picks[A], weight sums[A]].  ``sim_stats`` is BatchedRVOSimulator.stats (int64 [8])."""
    t = torch.where(agents_done.bool(), agents_time.double(), torch.full_like(agents_time.double(), max_time))
    world_tt = t.mean(1) + 3 * t.std(1, unbiased=False)
    head = torch.stack([
        torch.tensor(float(t.numel()), dtype=torch.float64, device=t.device),
        agents_done.double().sum(),
        sim_stats[2].double(), sim_stats[3].double(),
        torch.tensor(float(agent_steps), dtype=torch.float64, device=t.device),
        t.sum(), (t * t).sum(),
        torch.tensor(float(t.shape[0]), dtype=torch.float64, device=t.device),
        world_tt.sum(),
    ])
    if action_weights is None:
        return head
    A = action_weights.shape[-1]
    picks = torch.bincount(action_ids.reshape(-1).long(), minlength=A).double()
    wsum = action_weights.double().reshape(-1, A).sum(0)
    return torch.cat([head, picks, wsum])


def all_reduce_stats(packed: torch.Tensor, group=None) -> torch.Tensor:
    """The path's single collective: SUM over ranks, in place.  No-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def summarize(packed: torch.Tensor) -> EpisodeSummary:
    p = packed.detach().cpu().double()
    n = float(p[0])
    mean = float(p[5]) / n
    var = max(float(p[6]) / n - mean * mean, 0.0)
    std = var ** 0.5
    A = (p.numel() - 9) // 2
    picks = p[9:9 + A]
    wsum = p[9 + A:9 + 2 * A]
    return EpisodeSummary(agents=n, finished=float(p[1]), collisions=float(p[2]), lp3_calls=float(p[3]),
                          agent_steps=float(p[4]), mean_time=mean, std_time=std, ttime=mean + 3 * std,
                          mean_world_ttime=float(p[8]) / float(p[7]), action_picks=picks,
                          action_value=wsum / n if A else wsum)
