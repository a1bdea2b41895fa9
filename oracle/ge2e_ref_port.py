"""Torch-CPU port of the reference's *expanded* GE2E algorithm -- TEST/BENCH INFRASTRUCTURE.

Only ``tests/`` and ``bench.py`` (``cpu_baseline`` leg and ``--impl reference`` arm) may
import this.  It exists because ``/root/reference`` does not travel to the GPU box: this
port does the same arithmetic, in the same order and with the same O(N^2 M D) expanded
tensors, as ``embedding_model_GE2E/s3_loss_function_GE2E.py`` (cited ``s3:<line>``), so
that timing it on the box's host cores stands in for timing the reference itself.
``tests/test_oracle.py`` checks it bit-for-bit (fp32, same torch) against the fixtures
produced from the real class.

Unlike the reference it can also evaluate a *row sample*: the loss terms of a subset of
utterances against all N centroids, which costs (rows/U) of the full pass.  That is the
"bounded sample" used when the full batch cannot be expanded in host memory
(N=1024,M=10 needs ~10 GB per expanded tensor and >80 GB with autograd).
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F


def loo_centroids(E):
    """s3:95-112: (sum over the speaker's utterances - self) / (M - 1)."""
    total = E.sum(dim=1).reshape(E.shape[0], 1, E.shape[-1])
    return (total - E) / (E.shape[1] - 1)


def cos_sim_expanded(E, eps):
    """s3:41-80 with the reference's repeat-expansion (no matmul)."""
    N, M, D = E.shape
    C = E.mean(dim=1)                                              # s3:37
    Ef = E.view(N * M, D)
    same = F.cosine_similarity(Ef, loo_centroids(E).view(N * M, D))  # s3:57
    C_rep = C.repeat((N * M, 1))                                   # s3:64  [U*N, D]
    E_rep = Ef.unsqueeze(1).repeat(1, N, 1).view(N * M * N, D)     # s3:65-69
    cos = F.cosine_similarity(E_rep, C_rep).view(N, M, N)          # s3:70
    j = list(range(N))
    cos[j, :, j] = same.view(N, M)                                 # s3:77-78
    return cos + eps                                               # s3:79


def softmax_loss(S, eps):
    """s3:114-127."""
    j = list(range(S.size(0)))
    pos = S[j, :, j]
    neg = (torch.exp(S).sum(dim=2) + eps).log_()
    per = -1 * (pos - neg)
    return per.sum(), per


def loss_full(E, w, b, eps=1e-6):
    """s3:19-30: whole-batch loss, reference order of operations."""
    return softmax_loss(w * cos_sim_expanded(E, eps) + b, eps)[0]


def loss_row_sample(E, w, b, rows, eps=1e-6):
    """Same arithmetic restricted to the utterance rows ``rows`` (flat indices into U).

    Centroids still come from the whole batch; only the expanded cosine (the ~95 % cost,
    SURVEY.md section 2.1) is restricted.  Sum of the per-row losses of those rows.
    """
    N, M, D = E.shape
    C = E.mean(dim=1)
    Ef = E.view(N * M, D)[rows]
    Uf = loo_centroids(E).view(N * M, D)[rows]
    R = Ef.shape[0]
    same = F.cosine_similarity(Ef, Uf)
    C_rep = C.repeat((R, 1))
    E_rep = Ef.unsqueeze(1).repeat(1, N, 1).view(R * N, D)
    cos = F.cosine_similarity(E_rep, C_rep).view(R, N)
    spk = torch.as_tensor(rows, dtype=torch.long) // M
    cos[torch.arange(R), spk] = same
    S = w * (cos + eps) + b
    pos = S[torch.arange(R), spk]
    neg = (torch.exp(S).sum(dim=1) + eps).log_()
    return (neg - pos).sum()


def time_fwd_bwd(E_np, w=10.0, b=-5.0, eps=1e-6, rows=None, iters=1, warmup=0, threads=None):
    """Time loss + backward (grads to E, w, b) on the host; returns (seconds/iter, rows, loss)."""
    if threads:
        torch.set_num_threads(threads)
    E = torch.as_tensor(E_np, dtype=torch.float32).clone().requires_grad_(True)
    wt = torch.tensor(float(w), requires_grad=True)
    bt = torch.tensor(float(b), requires_grad=True)
    n_rows = E.shape[0] * E.shape[1] if rows is None else len(rows)
    last = None
    ts = []
    for it in range(warmup + iters):
        E.grad = wt.grad = bt.grad = None
        t0 = time.perf_counter()
        loss = loss_full(E, wt, bt, eps) if rows is None else loss_row_sample(E, wt, bt, rows, eps)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            ts.append(dt)
        last = float(loss.detach())
    ts.sort()
    return ts[len(ts) // 2], n_rows, last
