"""What bounds the host-fed e2e number: raw pinned H2D rate (one buffer vs rotating buffers, default
vs side stream, alone vs under the fwd+bwd graph) and GE2EHostFeed with 1 / 3 host batches."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speaker_embedding_ge2e_loss_b200 import GE2EHostFeed, GE2EPlan

dev = torch.device("cuda:0")
N, M, D = 1024, 10, 256
nb = N * M * D * 4
hosts = [torch.randn(N, M, D).pin_memory() for _ in range(8)]
dst = [torch.empty(N, M, D, device=dev) for _ in range(2)]

def rate(nh, stream, steps=200):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for k in range(10):
            dst[k % 2].copy_(hosts[k % nh], non_blocking=True)
        e0.record(stream)
        for k in range(steps):
            dst[k % 2].copy_(hosts[k % nh], non_blocking=True)
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3

side = torch.cuda.Stream()
cur = torch.cuda.current_stream()
for nh in (1, 2, 3, 8):
    us = rate(nh, side)
    print(f"H2D alone, side stream, {nh} host buffers: {us:.1f} us  {nb / us / 1e3:.1f} GB/s")
us = rate(1, cur)
print(f"H2D alone, default stream, 1 host buffer: {us:.1f} us  {nb / us / 1e3:.1f} GB/s")

# H2D on the side stream while the compute stream replays the fwd+bwd graph back to back
w = torch.tensor(10.0, device=dev); b = torch.tensor(-5.0, device=dev)
plan = GE2EPlan(N, M, D, "softmax", "tf32", device=dev)
E = torch.nn.functional.normalize(torch.randn(N, M, D, device=dev), dim=-1)
g = plan.capture(E, w, b, steps=10)
comp = torch.cuda.Stream()
for nh in (1, 3):
    torch.cuda.synchronize()
    with torch.cuda.stream(comp):
        for _ in range(60):
            g.replay()                 # ~0.7 ms each -> ~42 ms of compute
    us = rate(nh, side, steps=100)
    torch.cuda.synchronize()
    print(f"H2D under back-to-back fwd+bwd, {nh} host buffers: {us:.1f} us  {nb / us / 1e3:.1f} GB/s")

for nh in (1, 3, 1, 3):
    feed = GE2EHostFeed(N, M, D, w, b, "softmax", "tf32", device=dev)
    hs = [torch.nn.functional.normalize(torch.randn(N, M, D), dim=-1).pin_memory() for _ in range(nh)]
    def fed(steps):
        prev = None
        for k in range(steps):
            t = feed.submit(hs[k % nh])
            if prev is not None:
                feed.result(prev)
            prev = t
        feed.result(prev)
    def fed2(steps):
        for k in range(steps):
            t = feed.submit(hs[k % nh])
        feed.result(t)
    fed(5)
    torch.cuda.synchronize()
    line = []
    for rep in range(6):
        t0 = time.perf_counter(); fed(50); torch.cuda.synchronize()
        line.append((time.perf_counter() - t0) / 50 * 1e6)
    print(f"GE2EHostFeed, {nh} host batches, 6 x 50 steps: " + " ".join(f"{x:.0f}" for x in line) + " us/step")
    t0 = time.perf_counter(); fed2(200); torch.cuda.synchronize()
    us = (time.perf_counter() - t0) / 200 * 1e6
    print(f"GE2EHostFeed submit-only, {nh} host batches: {us:.1f} us/step")
