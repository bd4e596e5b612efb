#!/usr/bin/env python
"""PCIe: one direction at a time against both at once (pinned memory, two streams)."""
import time
import torch
dev = torch.device("cuda", 0)
n = 2_000_000_000
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
def chunked(k):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d1[i*c:(i+1)*c].copy_(h1[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h2[i*c:(i+1)*c].copy_(d2[i*c:(i+1)*c], non_blocking=True)
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D {n/a/1e9:.1f} GB/s ({a*1e3:.1f} ms)  D2H {n/b/1e9:.1f} GB/s ({b*1e3:.1f} ms)  both at once {2*n/c/1e9:.1f} GB/s total ({c*1e3:.1f} ms; serial would be {(a+b)*1e3:.1f})")
for k in (4, 16):
    x = t(lambda: chunked(k))
    print(f"both, {k} chunks each: {x*1e3:.1f} ms")
