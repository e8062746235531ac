"""N > 1 host logic on CPU: world_size-2 gloo process group.  Checks the env sharding and that
the packed statistics all-reduce reproduces the single-process numbers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from collision_avoidance_b200 import dist as cdist


def test_shard_ranges_partition_the_batch():
    for E in (1, 7, 64, 65536, 262144):
        for W in (1, 2, 3, 8):
            spans = [cdist.shard_range(E, r, W) for r in range(W)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == E
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _fake_rank_data(E, N, A, seed):
    rng = np.random.default_rng(seed)
    done = (rng.random((E, N)) < 0.7).astype(np.uint8)
    times = rng.uniform(1.0, 30.0, (E, N)).astype(np.float32)
    stats = rng.integers(0, 1000, 8).astype(np.int64)
    w = rng.uniform(-1, 1, (E, N, A)).astype(np.float32)
    ids = rng.integers(0, A, (E, N)).astype(np.uint8)
    return done, times, stats, w, ids


def _worker(rank, world, port, E, N, A, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = cdist.shard_range(E, rank, world)
        done, times, stats, w, ids = _fake_rank_data(E, N, A, 0)  # same global data on every rank
        sl = slice(start, start + count)
        # each rank owns a slice of the worlds; sim counters are split evenly for the test
        packed = cdist.pack_stats(torch.from_numpy(times[sl]), torch.from_numpy(done[sl]), 40.0,
                                  torch.from_numpy(stats) if rank == 0 else torch.zeros(8, dtype=torch.int64),
                                  agent_steps=count * N * 10, action_weights=torch.from_numpy(w[sl]),
                                  action_ids=torch.from_numpy(ids[sl]))
        cdist.all_reduce_stats(packed)
        q.put((rank, packed.numpy().copy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(120)
def test_stats_all_reduce_world_size_2_gloo():
    E, N, A, world = 10, 6, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, E, N, A, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # single-process reference over the whole batch
    done, times, stats, w, ids = _fake_rank_data(E, N, A, 0)
    ref = cdist.pack_stats(torch.from_numpy(times), torch.from_numpy(done), 40.0, torch.from_numpy(stats),
                           agent_steps=E * N * 10, action_weights=torch.from_numpy(w), action_ids=torch.from_numpy(ids))
    for r in range(world):
        np.testing.assert_allclose(results[r], ref.numpy(), rtol=1e-12, atol=1e-9)
    s = cdist.summarize(torch.from_numpy(results[0]))
    t = np.where(done == 1, times.astype(np.float64), 40.0)
    assert s.ttime == pytest.approx(t.mean() + 3 * t.std(), rel=1e-9)
    assert s.mean_world_ttime == pytest.approx((t.mean(1) + 3 * t.std(1)).mean(), rel=1e-9)
    assert s.finished == done.sum() and s.agents == E * N
    assert s.action_picks.sum().item() == E * N


def test_all_reduce_is_noop_without_process_group():
    x = torch.arange(5, dtype=torch.float64)
    assert torch.equal(cdist.all_reduce_stats(x.clone()), x)


def test_numa_binding_is_a_noop_without_a_gpu():
    """bind_to_gpu_numa never raises: without NVML / a GPU it reports None and leaves the affinity alone."""
    import os
    from collision_avoidance_b200.dist import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    cores = bind_to_gpu_numa(0)
    assert cores is None or set(cores) <= before
    if cores is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
