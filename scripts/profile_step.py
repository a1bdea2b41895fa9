"""Run a few bare fwd+bwd steps (no CUDA graph, no extra torch kernels) for ncu.
    python scripts/profile_step.py [cfg3|cfg2|cfg4] [tf32|fp32] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_batch  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import GE2EPlan  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N, M, D = WORKLOADS[wl]
dev = torch.device("cuda:0")
E = make_batch(N, M, D).to(dev)
w = torch.tensor(10.0, device=dev)
b = torch.tensor(-5.0, device=dev)
plan = GE2EPlan(N, M, D, "softmax", prec, device=dev)
for _ in range(steps):
    plan.step(E, w, b)
torch.cuda.synchronize()
print("loss", plan.loss.item(), "dw", plan.dw.item(), "path", plan.path)
