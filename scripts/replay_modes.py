"""Experiment: how should K steps be replayed to measure throughput?  (a) one graph per step, replayed
back to back; (b) one graph with K steps, cold first replay; (c) the same graph, second replay."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_batch  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import GE2EPlan  # noqa: E402

WL = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
N, M, D = WORKLOADS[WL]
dev = torch.device("cuda:0")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 50
NB = max(2, int(np.ceil(1.5 * 126e6 / (N * M * D * 4))))
batches = [make_batch(N, M, D, seed=i).to(dev) for i in range(NB)]
w = torch.tensor(10.0, device=dev)
b = torch.tensor(-5.0, device=dev)
plan = GE2EPlan(N, M, D, "softmax", "tf32", device=dev)


def timed(fn):
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) * 1e3 / K


graphs = [plan.capture(E, w, b) for E in batches]
for k in range(5):
    graphs[k % NB].replay()
print("(a) one graph per step, back to back: %.1f us/step" % timed(lambda: [graphs[k % NB].replay() for k in range(K)]))
print("(a) again                           : %.1f us/step" % timed(lambda: [graphs[k % NB].replay() for k in range(K)]))
g = plan.capture(batches, w, b, steps=K)
print("(b) K-step graph, first replay      : %.1f us/step" % timed(g.replay))
print("(c) K-step graph, second replay     : %.1f us/step" % timed(g.replay))
print("(c) K-step graph, third replay      : %.1f us/step" % timed(g.replay))
g5 = plan.capture(batches, w, b, steps=5)
print("(d) 5-step graphs x K/5             : %.1f us/step" % timed(lambda: [g5.replay() for _ in range(K // 5)]))
