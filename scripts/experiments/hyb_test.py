import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import speaker_embedding_ge2e_loss_b200 as pkg
from oracle import ge2e_oracle_torch as orct
dev = torch.device("cuda:0")
h = pkg.lib()
def run(N, M, D, mode, steps):
    h.ge2e_b200_debug_hybrid(mode)
    plan = pkg.GE2EPlan(N, M, D, "softmax", "tf32", device=dev)
    Es = [torch.nn.functional.normalize(torch.randn(N, M, D, device=dev, generator=torch.Generator(device=dev).manual_seed(i)), dim=-1) for i in range(4 if N < 4096 else 2)]
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
    dwerr = abs(plan.dw.item() - ref["dw"]) / max(1, abs(ref["dw"]))
    h.ge2e_b200_debug_hybrid(0)
    return float(np.median(ts)), plan.launches_per_step, err, lerr, dwerr
for (N, M, D, steps) in [(300, 7, 128, 20), (1024, 10, 256, 40), (2048, 16, 256, 10), (4096, 10, 256, 10), (8192, 2, 256, 8), (8192, 16, 256, 4)]:
    for mode in (-1, 1):
        t, nl, err, lerr, dwerr = run(N, M, D, mode, steps)
        print(N, M, D, "hybrid" if mode > 0 else "tf32  ", "%.1f us" % t, "launches", nl, "dE err %.2e loss %.1e dw %.1e" % (err, lerr, dwerr), flush=True)
