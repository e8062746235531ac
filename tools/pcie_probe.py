#!/usr/bin/env python
"""Host<->device copy bandwidth of the box with pinned buffers (dev tool): what bounds `e2e`."""
import torch

n = 16 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(2 * n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t = timed(lambda: d_in.copy_(h_in, non_blocking=True))
print(f"H2D 16 MiB: {t:.3f} ms  {n / t / 1e6:.1f} GB/s")
t = timed(lambda: h_out.copy_(d_out, non_blocking=True))
print(f"D2H 32 MiB: {t:.3f} ms  {2 * n / t / 1e6:.1f} GB/s")


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    cur.wait_stream(s1)
    cur.wait_stream(s2)


t = timed(both)
print(f"H2D 16 MiB || D2H 32 MiB: {t:.3f} ms")
for mb in (1, 2, 4):
    k = mb << 20
    t = timed(lambda: h_out[:k].copy_(d_out[:k], non_blocking=True), reps=50)
    print(f"D2H {mb} MiB: {t * 1e3:.1f} us  {k / t / 1e6:.1f} GB/s")
