"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's EER sweep.

Follows /root/reference/embedding_model_GE2E/s5_eval_model.py:
  * :57        thresholds 0.01 * i + 0.5, i in range(50) (Python floats);
  * :59        ``S_thres = S > thres`` -- S is the float32 similarity matrix [N, M, N]
               (``sim_matrix.detach().cpu().numpy()``, :46); numpy compares a float32 array with a
               Python float in float32, so the threshold is rounded to float32 first;
  * :80-82     FAR = sum_i (count(S_thres[i]) - count(S_thres[i, :, i])) / ((N - 1) / M / N);
  * :87-89     FRR = sum_i (M - count(S_thres[i][:, i])) / (M / N);
  * :92-97     keep the threshold with the smallest |FAR - FRR| (strict <, first one wins; diff
               starts at 1), EER = (FAR + FRR) / 2.
The denominators are the reference's own (they are not the population sizes; FAR and FRR are not
ratios in [0, 1]) and are reproduced as they are.  Pinned against the reference's own source text
executed on seeded matrices: tests/golden/make_eer_golden.py -> tests/golden/eer_reference_vectors.npz.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np


def default_thresholds():
    """s5:57."""
    return [0.01 * i + 0.5 for i in range(50)]


def threshold_counts(S, thresholds):
    """Integer accept counts per threshold: over the whole matrix and over the own-speaker entries
    S[i, :, i] (s5:59, :81, :88)."""
    S = np.asarray(S, dtype=np.float32)
    N, M, N2 = S.shape
    assert N == N2
    own = S[np.arange(N), :, np.arange(N)]                  # [N, M]
    acc_all = np.empty(len(thresholds), dtype=np.int64)
    acc_own = np.empty(len(thresholds), dtype=np.int64)
    for t, th in enumerate(thresholds):
        th32 = np.float32(th)
        acc_all[t] = int(np.count_nonzero(S > th32))
        acc_own[t] = int(np.count_nonzero(own > th32))
    return acc_all, acc_own


def eer_from_counts(acc_all, acc_own, N, M, thresholds):
    """s5:50-97 from the integer counts.  Returns dict(EER, thres, FAR, FRR, far[], frr[])."""
    diff, EER, EER_thres, EER_FAR, EER_FRR = 1, 0, 0, 0, 0
    fars, frrs = [], []
    for t, thres in enumerate(thresholds):
        denominator = (N - 1) / M / N                        # s5:80
        FAR = (int(acc_all[t]) - int(acc_own[t])) / denominator if denominator != 0 else float("inf")
        denominator = M / N                                  # s5:87
        FRR = (N * M - int(acc_own[t])) / denominator
        fars.append(FAR)
        frrs.append(FRR)
        if diff > abs(FAR - FRR):                            # s5:92
            diff = abs(FAR - FRR)
            EER = (FAR + FRR) / 2
            EER_thres = thres
            EER_FAR = FAR
            EER_FRR = FRR
    return dict(EER=EER, thres=EER_thres, FAR=EER_FAR, FRR=EER_FRR, far=fars, frr=frrs)


def eer_sweep(S, thresholds=None):
    thresholds = default_thresholds() if thresholds is None else list(thresholds)
    S = np.asarray(S, dtype=np.float32)
    a, o = threshold_counts(S, thresholds)
    out = eer_from_counts(a, o, S.shape[0], S.shape[1], thresholds)
    out["accept_all"], out["accept_own"] = a, o
    return out
