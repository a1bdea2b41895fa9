"""what folding the trainer's `embeddings[unperm]` gather (and the scatter of its
backward) into the loss kernels saves.  Eager module API, fwd + bwd, device time per call (CUDA
events around K calls) and host time per call.
    python scripts/unperm_cost.py [cfg3|cfg2|cfg4] [tf32|fp32]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_batch  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import GE2ELoss  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
N, M, D = WORKLOADS[wl]
dev = torch.device("cuda:0")
crit = GE2ELoss(None, device=dev, precision=prec)
perm = torch.randperm(N * M, device=dev)
unperm = torch.argsort(perm).to(torch.int32)
flat = make_batch(N, M, D).to(dev).reshape(N * M, D)[perm].contiguous()


def run(fused, K=200, W=20):
    x = flat.clone().requires_grad_(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for k in range(W + K):
        if k == W:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ev[0].record()
        x.grad = None
        loss = crit(x, unperm=unperm, speakers=N) if fused else crit(x[unperm.long()].reshape(N, M, D))
        loss.backward()
    ev[1].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) * 1e3 / K, (t1 - t0) * 1e6 / K, loss.item()


a = run(False)
b = run(True)
print(f"{wl} {prec}: gather in torch  : {a[0]:7.1f} us device, {a[1]:7.1f} us host per fwd+bwd  (loss {a[2]:.4f})")
print(f"{wl} {prec}: gather in kernels: {b[0]:7.1f} us device, {b[1]:7.1f} us host per fwd+bwd  (loss {b[2]:.4f})")
