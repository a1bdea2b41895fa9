import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import speaker_embedding_ge2e_loss_b200 as pkg
from oracle import ge2e_oracle_torch as orct
dev = torch.device("cuda:0")
def timeit(N, M, D, prec):
    plan = pkg.GE2EPlan(N, M, D, "softmax", prec, device=dev)
    Es = [torch.nn.functional.normalize(torch.randn(N, M, D, device=dev), dim=-1) for _ in range(8)]
    w = torch.tensor(10.0, device=dev); b = torch.tensor(-5.0, device=dev)
    g = plan.capture(Es, w, b, steps=40)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) / 40 * 1e3)
    plan.step(Es[0], w, b); torch.cuda.synchronize()
    ref = orct.forward_backward(Es[0], 10.0, -5.0, 1e-6, "softmax")
    err = ((plan.dE.double() - ref["dE"]).norm() / ref["dE"].norm()).item()
    return float(np.median(ts)), plan.path, plan.single_kernel, err
for (N, M, D) in [(129, 10, 256), (160, 10, 256), (200, 10, 256), (255, 10, 256), (256, 10, 256), (200, 4, 128)]:
    for prec in ("fp32_simt", "fp32", "tf32"):
        t, path, sk, err = timeit(N, M, D, prec)
        print(N, M, D, prec, "%.1f us" % t, "path", path, "single", sk, "dE err %.2e" % err)
