"""Stage timeline of the single-kernel small-batch step (CTA 0, SM clocks): GE2E_SMALL_STOP=99 python scripts/small_step_trace.py [cfg2]"""
import os
import sys

os.environ["GE2E_SMALL_STOP"] = "99"
os.environ.setdefault("GE2E_SMALL_STEP", "2")        # trace every supported shape, not only the selected ones
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_batch  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import GE2EPlan  # noqa: E402

dev = torch.device("cuda:0")
NAMES = ["start", "prep", "barrier1", "load c_hat/e_hat", "cos block", "softmax rows", "dE_hat rows", "dC shares",
         "barrier2", "dC gather", "finalize"]
for wl in (sys.argv[1:] or ["cfg2"]):
    N, M, D = WORKLOADS[wl]
    E = make_batch(N, M, D).to(dev)
    w = torch.tensor(10.0, device=dev); b = torch.tensor(-5.0, device=dev)
    plan = GE2EPlan(N, M, D, "softmax", "fp32", device=dev)
    for _ in range(300):
        plan.step(E, w, b)
    torch.cuda.synchronize()
    st = plan.row_kstar[:11].cpu().numpy().astype("int64") & 0xffffffff
    print(wl, "loss", plan.loss.item())
    for i in range(1, 11):
        print(f"  {NAMES[i]:18s} {((st[i] - st[i - 1]) & 0xffffffff):8d} clk")
    print(f"  {'total':18s} {((st[10] - st[0]) & 0xffffffff):8d} clk")
