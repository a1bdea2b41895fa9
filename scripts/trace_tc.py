"""Pipeline timeline of the tensor-core kernels from the library's debug trace.
    python scripts/trace_tc.py [cfg3] [fwd|bwd]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_batch  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import _lib, lib, ops  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
N, M, D = WORKLOADS[wl]
dev = torch.device("cuda:0")
E = make_batch(N, M, D).to(dev)
w = torch.tensor(10.0, device=dev)
b = torch.tensor(-5.0, device=dev)
g = torch.tensor(1.0, device=dev)
c_hat = torch.empty((N, D), device=dev)
e_hat, cos_diag, accum = ops.prep(E, c_hat, _lib.TF32)
G = 148  # grid of the default cluster size 2 (74 clusters)
trace = torch.zeros((G, 3, 64), dtype=torch.int64, device=dev)


def show(name, tr):
    tr = tr.cpu().numpy().astype(np.int64)
    used = tr[:, :, 0] > 0
    t0 = tr[tr > 0].min()
    rel = np.where(tr > 0, tr - t0, -1)
    print(f"== {name}: CTAs used {used[:, 1].sum()}, span {(tr.max() - t0) / 1e3:.1f} us")
    for role, rn in enumerate(("tma", "mma", "epi")):
        ends = np.array([rel[c, role][rel[c, role] >= 0].max() for c in range(G) if used[c, role]])
        starts = np.array([rel[c, role][0] for c in range(G) if used[c, role]])
        print(f"  {rn}: first-stamp min/med/max = {starts.min()/1e3:.2f}/{np.median(starts)/1e3:.2f}/{starts.max()/1e3:.2f} us"
              f"  last-stamp min/med/max = {ends.min()/1e3:.2f}/{np.median(ends)/1e3:.2f}/{ends.max()/1e3:.2f} us")
    for c in (0, 1, 73, 147):
        if not used[c, 1]:
            continue
        for role, rn in enumerate(("tma", "mma", "epi")):
            v = rel[c, role]
            v = v[v >= 0]
            print(f"  cta{c:3d} {rn}: " + " ".join(f"{x/1e3:.2f}" for x in v))


for rep in range(2):
    trace.zero_()
    lib().ge2e_b200_debug_trace(trace.data_ptr(), 0)
    accum.zero_()
    rs, ks, aux, per, _ = ops.fwd_rows(e_hat, c_hat, cos_diag, N, N, 0, M, D, w, b, 1e-6, 0, _lib.TF32, accum)
    torch.cuda.synchronize()
    lib().ge2e_b200_debug_trace(None, -1)
    if rep == 1:
        show("fwd (warm)", trace)
for mode, name in ((1, "bwd (dC then dE segments)"),):
    for rep in range(2):
        trace.zero_()
        lib().ge2e_b200_debug_trace(trace.data_ptr(), mode)
        out = ops.bwd_rows(e_hat, c_hat, cos_diag, rs, ks, aux, N, N, 0, M, D, w, b, 1e-6, 0, _lib.TF32, g)
        torch.cuda.synchronize()
        lib().ge2e_b200_debug_trace(None, -1)
        if rep == 1:
            show(name + " (warm)", trace)
