#!/usr/bin/env python
"""Headline benchmark: ORCA agent-steps/sec (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at every N: BASELINE.json configs[1] per GPU -- circle-crossing, 16 agents/env x
65,536 envs, ORCA policy (doStep + goal-directed preferred velocity + done test, i.e. the
reference's run_sim(mode=0) loop, ALAN_true.py:106-131,631-636,547-566).  Env instances are
independent, so N GPUs each own 65,536 envs (weak scaling, no data-path collective); episode
statistics are all-reduced once after the timed region.

A "step" is one fused environment step over every env of the rank (one kernel launch).
  value      device-timed throughput, state resident in HBM (CUDA events around each launch,
             L2 flushed between launches)
  e2e        same metric through the host-buffer C-ABI call (orca_step_host): goals go
             host->device and positions+velocities come device->host inside the timed region
  roofline   algorithmic HBM bytes / kernel time vs. the measured copy bandwidth
  cpu_baseline  the CPU oracle (restatement of RVO2, oracle/) on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "orca_agent_steps_per_sec"
UNIT = "agent-steps/s"
WORKLOAD = "circle-crossing 16 agents/env x 65,536 envs, ORCA policy, per B200"
ENVS_PER_GPU = 65536
AGENTS = 16
# algorithmic HBM bytes per agent-step of the fused step (DESIGN.md "Roofline"):
# read pos 8 + vel 8 + goal 8 + done flag 1, write pos 8 + vel 8
BYTES_PER_AGENT_STEP = 41
SEED = 1234


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst)"
    return 6650.0, "fallback (B200_PROFILING.md)"


_RESULT_OUT = sys.stdout  # replaced in main() by a private duplicate of the real stdout


def _traffic():
    """dram bytes per launch of the step kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("step_small_kernel_dram_bytes_per_launch")
    return None


def _issue_profile():
    """What actually bounds the step kernel (same ncu capture): issue-slot use, warp-instructions
    per launch and live lanes per instruction.  Reported next to the HBM roofline, which this
    kernel cannot approach."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    if "step_small_kernel_issue_active_pct" not in t:
        return None
    return {"issue_slots_busy_frac": t["step_small_kernel_issue_active_pct"] / 100.0,
            "warp_instructions_per_launch": t["step_small_kernel_warp_instructions_per_launch"],
            "active_lanes_per_instruction": t["step_small_kernel_active_lanes_per_instruction"],
            "source": "profiles/r01_step_small_ncu_raw.csv (ncu --set full, bench step ~100)"}


class ClockSampler:
    """Samples SM clocks + throttle reasons (NVML, every few ms) while the timed region runs."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []  # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES remapping
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                self.samples.append((mhz, reasons))
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self._nvml is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        nv = self._nvml
        sm = sorted(s[0] for s in self.samples)
        bits = 0
        for s in self.samples:
            bits |= s[1]
        names = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")]
        alt = {"nvmlClocksEventReasonHwSlowdown": 0x8, "nvmlClocksEventReasonHwThermalSlowdown": 0x40,
               "nvmlClocksEventReasonSwThermalSlowdown": 0x20, "nvmlClocksEventReasonSwPowerCap": 0x4}
        reasons = [n for n, attr in names if bits & int(getattr(nv, attr, alt[attr]))]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
def cpu_baseline(steps: int, warmup: int, target_seconds: float = 12.0, threads: int | None = None):
    """Times the CPU oracle (the restatement of RVO2 the reference calls into) on the same
    scenario and the same step window as the GPU run, one worker thread per host core.
    Bounded sample: fewer envs, sized for ~target_seconds of CPU work."""
    from collision_avoidance_b200 import scenarios
    from oracle import rvo2_oracle
    from oracle.helpers import oracle_sims
    threads = threads or (os.cpu_count() or 1)
    # calibrate on a small batch over the same window, then size the sample
    cal_envs = 4 * threads
    scn = scenarios.circle(cal_envs, AGENTS, seed=SEED)
    sims = oracle_sims(scn)
    goals = scn.goal.astype("float64")
    t0 = time.perf_counter()
    rvo2_oracle.batch_orca_steps(sims, goals, warmup + steps, threads)
    cal = time.perf_counter() - t0
    envs = int(max(cal_envs, min(65536, cal_envs * target_seconds / max(cal, 1e-6))))
    envs -= envs % threads
    scn = scenarios.circle(envs, AGENTS, seed=SEED)
    sims = oracle_sims(scn)
    goals = scn.goal.astype("float64")
    rvo2_oracle.batch_orca_steps(sims, goals, warmup, threads)
    t0 = time.perf_counter()
    rvo2_oracle.batch_orca_steps(sims, goals, steps, threads)
    dt = time.perf_counter() - t0
    return {
        "value": envs * AGENTS * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"{envs} envs x {AGENTS} agents x {steps} steps after {warmup} warm-up steps of the same circle "
                  f"scenario (C++ oracle: kd-tree + ORCA + LP, float64 pref update), {dt:.1f} s wall",
    }


def cpu_python_driver_baseline(steps: int = 30, envs: int = 8):
    """The reference's own call pattern on one core: per agent setAgentPrefVelocity, one doStep,
    per agent getAgentPosition/getAgentVelocity, all through Python (collision_avoidence_env.py
    :371-400), with the oracle standing in for rvo2.  This is what "Python-RVO2 as the reference
    uses it" costs; reported next to the C++-driver figure."""
    from collision_avoidance_b200 import scenarios
    from oracle.shell_oracle import AlanShellOracle
    scn = scenarios.circle(envs, AGENTS, seed=SEED)
    shells = [AlanShellOracle(scn, e) for e in range(envs)]
    for sh in shells:
        sh.orca_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        for sh in shells:
            sh.orca_step()
            sh.step_count += 1
            sh.done_test()
    dt = time.perf_counter() - t0
    return {"value": envs * AGENTS * steps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{envs} envs x {AGENTS} agents x {steps} steps, Python per-agent call pattern over the oracle"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  rvo2
    (Python-RVO2) is not available offline, so this times the oracle port of it (oracle/)
    with all host threads on the same config; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from collision_avoidance_b200 import scenarios
    from oracle import rvo2_oracle
    from oracle.helpers import oracle_sims
    threads = os.cpu_count() or 1
    envs = 2048
    scn = scenarios.circle(envs, AGENTS, seed=SEED)
    sims = oracle_sims(scn)
    goals = scn.goal.astype("float64")
    rvo2_oracle.batch_orca_steps(sims, goals, args.warmup, threads)
    t0 = time.perf_counter()
    rvo2_oracle.batch_orca_steps(sims, goals, args.steps, threads)
    dt = time.perf_counter() - t0
    value = envs * AGENTS * args.steps / dt
    sample = f"{envs} envs x {AGENTS} agents per step (bounded sample of the 65,536-env workload), {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_timed": envs, "agents_per_env": AGENTS,
                   "note": "rvo2 (Python-RVO2) unavailable offline; oracle port of RVO2 timed on host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from collision_avoidance_b200 import _lib, scenarios
    from collision_avoidance_b200.sim import BatchedRVOSimulator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: run on (and pin host buffers from) the cores next to that GPU
    from collision_avoidance_b200.dist import bind_to_gpu_numa
    # (only with several ranks: at N=1 the CPU baseline below wants every host core)
    cores = bind_to_gpu_numa(local_rank) if world > 1 and not os.environ.get("BENCH_NO_AFFINITY") else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    E, N = ENVS_PER_GPU, AGENTS
    scn = scenarios.circle(E, N, seed=SEED + rank)
    sim = BatchedRVOSimulator(E, N, device=dev, **scn.params)
    sim.set_obstacles(scn.obstacles)

    def reset_state():
        sim.pos.copy_(torch.from_numpy(scn.pos))
        sim.vel.copy_(torch.from_numpy(scn.vel))
        st = dict(goal=torch.from_numpy(scn.goal).to(dev), goal2=torch.from_numpy(scn.goal2).to(dev),
                  agent_done=torch.zeros(E, N, dtype=torch.uint8, device=dev),
                  arrival_time=torch.full((E, N), 0.0, device=dev),
                  env_step=torch.zeros(E, dtype=torch.int32, device=dev),
                  env_done_cnt=torch.zeros(E, dtype=torch.int32, device=dev))
        sim.stats.zero_()
        return st

    st = reset_state()

    def step():
        sim.env_step(policy=_lib.POLICY_GOAL, done_mode=_lib.DONE_GOAL_RADIUS, **st)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed region -----------------------------------------------------------
    for _ in range(args.warmup):
        step()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = sim.launch_count()
    barrier()
    with ClockSampler(local_rank) as clocks:
        wall0 = time.perf_counter()
        for i in range(args.steps):
            flush_buf.zero_()  # evict the state from L2 between timed launches
            starts[i].record()
            step()
            ends[i].record()
        barrier()
        wall = time.perf_counter() - wall0
    launches = sim.launch_count() - launches0
    kernel_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = sum(kernel_ms)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * E * N * args.steps / (total_ms_max * 1e-3)

    # episode statistics: the only collective of the path (one packed all-reduce, after timing)
    stats = sim.stats.clone()
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    stats_d = {"finished": int(stats[_lib.STAT_FINISHED]), "collisions": int(stats[_lib.STAT_COLLISIONS]),
               "lp3_calls": int(stats[_lib.STAT_LP3_CALLS]), "overflow": int(stats[_lib.STAT_OVERFLOW])}

    # ---- end-to-end region: host buffers through orca_step_host ---------------------------
    pos_h = torch.from_numpy(scn.pos.copy()).pin_memory()
    vel_h = torch.from_numpy(scn.vel.copy()).pin_memory()
    goal_h = torch.from_numpy(scn.goal.copy()).pin_memory()
    # same episode phase as the kernel-timed region: W steps from the reset state, then K timed
    e2e_steps = max(3, args.steps)
    sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=True, steps=max(1, args.warmup))
    for _ in range(8):  # untimed: the library times both of its routes (direct / staged) on its first six steady calls and keeps the faster
        sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        # per step: goals host->device (the setAgentPrefVelocity traffic), doStep, then
        # positions + velocities device->host (the getAgentPosition/Velocity traffic)
        sim.step_host(pos_h, vel_h, goal_h, policy=_lib.POLICY_GOAL, upload_state=False, steps=1)
    barrier()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * E * N * e2e_steps / float(te.item())
    bytes_state = E * N * 8

    if rank == 0:
        peak, peak_src = _peaks()
        avg_kernel_s = (total_ms / args.steps) * 1e-3  # rank 0's own launches
        achieved = BYTES_PER_AGENT_STEP * E * N / avg_kernel_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E, "agents_per_env": N, "policy": "orca_step + done_test",
                       "sharding": f"envs x{world}, no data-path collective", "l2": "flushed between timed launches "
                       "(256 MiB memset; state is 41 MB/GPU, smaller than L2)", "timing": "CUDA events per launch, "
                       "summed; max over ranks"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(), "peak_source": peak_src,
                         "bytes_per_agent_step": BYTES_PER_AGENT_STEP, "kernel": "step_small_kernel<10,GOAL>",
                         "note": "kernel is instruction-issue / latency bound, not HBM bound; see DESIGN.md Roofline",
                         "issue_profile": _issue_profile()},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_state,
                    "d2h_bytes_per_step": 2 * bytes_state, "steps": e2e_steps,
                    "api": "BatchedRVOSimulator.step_host -> orca_step_host (pinned host buffers; the library keeps the faster of its two routes: kernel reads/writes the mapped host buffers directly, or chunked upload | step | download over streams)"},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "host_affinity": {"rank0_cores": len(cores) if cores else None,
                              "note": "each rank runs on the cores NVML lists next to its GPU" if cores else "not bound"},
            "wall_s_timed_region": wall,
            "episode_stats": stats_d,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.steps, args.warmup)
            line["cpu_baseline_python_driver"] = cpu_python_driver_baseline()
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries may write there too (NCCL prints its
    # version banner to stdout when NCCL_DEBUG is set, as on the GPU boxes): keep a private handle on
    # the real stdout for the line and point fd 1 at stderr for everything else.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
