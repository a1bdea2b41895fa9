"""The reference's own algorithm (expanded repeat + F.cosine_similarity, oracle/ge2e_ref_port.py, a
line-by-line port of s3:19-127) run eagerly ON THE B200, beside this repo's loss on the same inputs:
how much of the GPU-vs-CPU ratio is the hardware and how much is the formulation.  cfg3 needs tens of
GB in that formulation (the [U*N, D] expansions and their autograd copies); it fits the 180 GB of HBM.
    python scripts/ref_on_gpu.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch  # noqa: E402
from oracle import ge2e_ref_port as port  # noqa: E402
import speaker_embedding_ge2e_loss_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
for (N, M, D, iters) in [(64, 10, 256, 10), (256, 10, 256, 5), (512, 10, 256, 3), (1024, 10, 256, 2)]:
    E0 = make_batch(N, M, D, seed=0).to(dev)
    res = {"N": N, "M": M, "D": D}
    try:
        torch.cuda.reset_peak_memory_stats()
        ts = []
        for it in range(iters + 1):
            E = E0.clone().requires_grad_(True)
            w = torch.tensor(10.0, device=dev, requires_grad=True)
            b = torch.tensor(-5.0, device=dev, requires_grad=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss = port.loss_full(E, w, b)
            loss.backward()
            e1.record()
            torch.cuda.synchronize()
            if it > 0:
                ts.append(e0.elapsed_time(e1))
            ref_loss, ref_dE = loss.item(), E.grad.clone()
            del loss, E
        res["reference_formulation_ms"] = float(np.median(ts))
        res["reference_formulation_peak_GB"] = torch.cuda.max_memory_allocated() / 1e9
    except torch.OutOfMemoryError as ex:
        res["reference_formulation"] = "out of memory: " + str(ex)[:80]
        ref_loss = ref_dE = None
    torch.cuda.empty_cache()
    for prec in ("fp32", "tf32"):
        plan = pkg.GE2EPlan(N, M, D, "softmax", prec, device=dev)
        w = torch.tensor(10.0, device=dev)
        b = torch.tensor(-5.0, device=dev)
        g = plan.capture(E0, w, b)
        for _ in range(5):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[f"this_repo_{prec}_ms"] = e0.elapsed_time(e1) / 50
        if ref_loss is not None:
            res[f"this_repo_{prec}_vs_ref_loss_rel"] = abs(plan.loss.item() - ref_loss) / abs(ref_loss)
            res[f"this_repo_{prec}_vs_ref_dE_rel"] = ((plan.dE - ref_dE).norm() / ref_dE.norm()).item()
    if "reference_formulation_ms" in res:
        res["speedup_tf32"] = res["reference_formulation_ms"] / res["this_repo_tf32_ms"]
        res["speedup_fp32"] = res["reference_formulation_ms"] / res["this_repo_fp32_ms"]
    print(json.dumps(res), flush=True)
    del ref_dE
    torch.cuda.empty_cache()
