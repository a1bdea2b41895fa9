import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import speaker_embedding_ge2e_loss_b200 as pkg
from oracle import ge2e_oracle_torch as orct
dev = torch.device("cuda:0")
def timeit(N, M, D, prec, steps):
    plan = pkg.GE2EPlan(N, M, D, "softmax", prec, device=dev)
    nb = max(2, int(200e6 // (N * M * D * 4)) + 1) if N * M * D * 4 < 100e6 else 2
    Es = [torch.nn.functional.normalize(torch.randn(N, M, D, device=dev), dim=-1) for _ in range(min(nb, 8))]
    w = torch.tensor(10.0, device=dev); b = torch.tensor(-5.0, device=dev)
    g = plan.capture(Es, w, b, steps=steps)
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) / steps * 1e3)
    plan.step(Es[0], w, b); torch.cuda.synchronize()
    ref = orct.forward_backward(Es[0], 10.0, -5.0, 1e-6, "softmax", chunk=2048)
    err = ((plan.dE.double() - ref["dE"]).norm() / ref["dE"].norm()).item()
    lerr = abs(plan.loss.item() - ref["loss"]) / abs(ref["loss"])
    return float(np.median(ts)), plan.path, err, lerr
for (N, M, D, steps) in [(1024, 10, 256, 40), (2048, 10, 256, 20), (4096, 10, 256, 10), (2048, 16, 256, 10), (8192, 16, 256, 4), (1024, 16, 256, 8), (8192, 2, 256, 8)]:
    for prec in ("tf32_mma", "f16"):
        t, path, err, lerr = timeit(N, M, D, prec, steps)
        print(N, M, D, "U*N=%.0fM" % (N * M * N / 1e6), prec, "%.1f us" % t, "path", path, "dE err %.2e loss err %.1e" % (err, lerr), flush=True)
