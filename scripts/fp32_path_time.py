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
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{wl} fp32 path: {us:.1f} us/step  {N * M / us:.2f} M utt/s  {6.0 * N * M * N * D / us / 1e6:.1f} TFLOP/s  loss {plan.loss.item():.4f}")
