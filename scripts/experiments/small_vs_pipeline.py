import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import speaker_embedding_ge2e_loss_b200 as pkg
dev = torch.device("cuda:0")
h = pkg.lib()
def timeit(N, M, D, variant, mode):
    h.ge2e_b200_debug_small_step(mode)
    plan = pkg.GE2EPlan(N, M, D, variant, "fp32_simt", device=dev)
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
    h.ge2e_b200_debug_small_step(1)
    return float(np.median(ts)), plan.single_kernel
for (N, M, D) in [(64, 10, 256), (96, 10, 256), (128, 10, 256), (128, 16, 256), (128, 4, 128), (32, 10, 256), (16, 16, 256)]:
    for variant in ("softmax", "contrast"):
        tp, _ = timeit(N, M, D, variant, 0)
        ts, sk = timeit(N, M, D, variant, 2)
        print(N, M, D, variant, "pipeline %.1f us" % tp, "single %.1f us" % ts, "(single kernel taken: %s)" % sk)
