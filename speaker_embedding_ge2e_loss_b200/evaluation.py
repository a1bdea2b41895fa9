"""EER sweep of the reference's evaluation script on the GPU.

Reference: embedding_model_GE2E/s5_eval_model.py:42-98 -- the similarity matrix of a test batch is
copied to the host and swept over 50 thresholds with numpy (`S > thres`, per-speaker sums), keeping the
threshold where FAR and FRR are closest.  Here the counting is one CUDA kernel over the device-resident
matrix for all thresholds at once (``ge2e_b200_threshold_counts``); what stays on the host is the
reference's scalar arithmetic on those integer counts (its own denominators and its own selection rule,
reproduced as they are, so the numbers printed by s5 are reproduced exactly).

    res = eer_sweep(sim_matrix)                 # sim_matrix [N, M, N] on the GPU
    res = evaluate_eer(embeddings)              # s5:42-46 + the sweep, from [N, M, D] embeddings
    print("EER : %0.2f (thres:%0.2f, FAR:%0.2f, FRR:%0.2f)" % (res.EER, res.thres, res.FAR, res.FRR))

CUDA (sm_100a) only; there is no host fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops


def default_thresholds() -> List[float]:
    """s5:57."""
    return [0.01 * i + 0.5 for i in range(50)]


@dataclass
class EERResult:
    EER: float
    thres: float
    FAR: float
    FRR: float
    thresholds: List[float] = field(repr=False)
    far: List[float] = field(repr=False)               # per threshold, the reference's FAR (s5:80-82)
    frr: List[float] = field(repr=False)               # per threshold, the reference's FRR (s5:87-89)
    accept_all: np.ndarray = field(repr=False)         # int64[T]: entries of S above the threshold
    accept_own: np.ndarray = field(repr=False)         # int64[T]: own-speaker entries above the threshold


def threshold_counts(sim_matrix: torch.Tensor, thresholds: Sequence[float]):
    """(accept_all, accept_own) as int64 numpy arrays in the order of ``thresholds``."""
    ops._need_cuda(sim_matrix)
    S = sim_matrix.detach()
    if S.dim() != 3 or S.shape[0] != S.shape[2]:
        raise ValueError(f"sim_matrix must be [N, M, N], got {tuple(S.shape)}")
    S = S.float().contiguous()
    N, M, _ = S.shape
    T = len(thresholds)
    if T < 1:
        raise ValueError("need at least one threshold")
    # numpy compares a float32 array with a Python float in float32 (s5:59): round the thresholds first
    th32 = np.asarray([np.float32(t) for t in thresholds], dtype=np.float32)
    order = np.argsort(th32, kind="stable")
    dev = S.device
    h = _lib.lib()
    with torch.cuda.device(dev):
        thr_dev = torch.from_numpy(th32[order].copy()).to(dev)
        counts = torch.empty((2, T), dtype=torch.int64, device=dev)
        nbytes = h.ge2e_b200_threshold_counts_scratch_bytes(T)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(h.ge2e_b200_threshold_counts(S.data_ptr(), N, M, thr_dev.data_ptr(), T, counts[0].data_ptr(),
                                                counts[1].data_ptr(), scratch.data_ptr(), nbytes, ops._stream()),
                   "ge2e_b200_threshold_counts")
        got = counts.cpu().numpy()
    acc_all = np.empty(T, dtype=np.int64)
    acc_own = np.empty(T, dtype=np.int64)
    acc_all[order] = got[0]
    acc_own[order] = got[1]
    return acc_all, acc_own


def eer_from_counts(acc_all, acc_own, N: int, M: int, thresholds: Sequence[float]) -> EERResult:
    """The reference's scalar arithmetic (s5:50-97) on the integer accept counts; host only."""
    if N < 2:
        raise ValueError("the reference's FAR denominator (N - 1) / M / N is zero for N = 1 (s5:80)")
    thresholds = list(thresholds)
    diff, EER, EER_thres, EER_FAR, EER_FRR = 1, 0, 0, 0, 0
    fars, frrs = [], []
    for t, thres in enumerate(thresholds):
        far = (int(acc_all[t]) - int(acc_own[t])) / ((N - 1) / M / N)      # s5:80-82
        frr = (N * M - int(acc_own[t])) / (M / N)                          # s5:87-89
        fars.append(far)
        frrs.append(frr)
        if diff > abs(far - frr):                                          # s5:92-97
            diff = abs(far - frr)
            EER, EER_thres, EER_FAR, EER_FRR = (far + frr) / 2, thres, far, frr
    return EERResult(EER, EER_thres, EER_FAR, EER_FRR, thresholds, fars, frrs, np.asarray(acc_all), np.asarray(acc_own))


def eer_sweep(sim_matrix: torch.Tensor, thresholds: Optional[Sequence[float]] = None) -> EERResult:
    """s5:50-97 on a device-resident similarity matrix: counts on the GPU, scalars on the host."""
    thresholds = default_thresholds() if thresholds is None else [float(t) for t in thresholds]
    N, M = int(sim_matrix.shape[0]), int(sim_matrix.shape[1])
    if N < 2:
        raise ValueError("the reference's FAR denominator (N - 1) / M / N is zero for N = 1 (s5:80)")
    acc_all, acc_own = threshold_counts(sim_matrix, thresholds)
    return eer_from_counts(acc_all, acc_own, N, M, thresholds)


def evaluate_eer(embeddings: torch.Tensor, w: float = 1.0, b: float = 0.0, hp=None,
                 thresholds: Optional[Sequence[float]] = None) -> EERResult:
    """s5:42-46 + the sweep: centroids, leave-one-out cosine matrix, ``w * cos + b`` (the script uses
    w = 1, b = 0), then ``eer_sweep`` -- the matrix never leaves the device."""
    from .loss import GE2ELoss
    cos = GE2ELoss.get_cos_sim(embeddings, GE2ELoss.get_centroids(embeddings), hp)
    S = cos if (w == 1.0 and b == 0.0) else w * cos + b
    return eer_sweep(S, thresholds)
