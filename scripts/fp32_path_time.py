"""fwd+bwd time of the precision="fp32" (SIMT) path, CUDA-graph replay: python scripts/fp32_path_time.py [cfg3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_batch  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import GE2EPlan  # noqa: E402

dev = torch.device("cuda:0")
for wl in (sys.argv[1:] or ["cfg2", "cfg3"]):
    N, M, D = WORKLOADS[wl]
    E = make_batch(N, M, D).to(dev)
    w = torch.tensor(10.0, device=dev)
    b = torch.tensor(-5.0, device=dev)
    plan = GE2EPlan(N, M, D, "softmax", "fp32", device=dev)
    g = plan.capture(E, w, b)
    import time
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.5:          # let the clocks ramp: these steps are tens of microseconds
        for _ in range(200):
            g.replay()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2000):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 2000 * 1e3
    print(f"{wl} fp32 path: {us:.1f} us/step  {N * M / us:.2f} M utt/s  {6.0 * N * M * N * D / us / 1e6:.1f} TFLOP/s  loss {plan.loss.item():.4f}")
