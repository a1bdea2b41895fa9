"""Host-side cost of one fwd+bwd through the module API (eager): wall time per step with the GPU
kept busy but never waited on inside the loop, plus a cProfile of the hot Python frames."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_batch  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import GE2ELoss  # noqa: E402

N, M, D = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
dev = torch.device("cuda:0")
crit = GE2ELoss(None, device=dev, precision="tf32")
E = make_batch(N, M, D).to(dev)


def step():
    Ed = E.detach().requires_grad_(True)
    loss = crit(Ed)
    crit.w.grad = crit.b.grad = None
    loss.backward()
    return loss


for _ in range(20):
    step()
torch.cuda.synchronize()
K = 300
t0 = time.perf_counter()
for _ in range(K):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per fwd+bwd: {(t1 - t0) / K * 1e6:.1f} us  (GPU drained {1e6 * (t2 - t1):.0f} us after the loop)")
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
