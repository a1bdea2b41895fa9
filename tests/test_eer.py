"""EER sweep (SURVEY 8(f) row 3, s5_eval_model.py:57-98).

CPU: the oracle restatement against vectors produced by the reference's own source text
(tests/golden/make_eer_golden.py).  GPU: the CUDA count kernel + host arithmetic against the oracle,
bit-exact (integer counts; FAR / FRR / EER are then the same float operations on the same integers).
"""
import os

import numpy as np
import pytest
import torch

from oracle import eer_oracle as eo
from oracle import ge2e_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eer_reference_vectors.npz")


def golden_cases():
    z = np.load(GOLD)
    return [(k[:-2], z[k], z[k[:-2] + "_res"], z[k[:-2] + "_far"], z[k[:-2] + "_frr"]) for k in z.files if k.endswith("_S")]


CASES = golden_cases()


@pytest.mark.parametrize("name,S,res,far,frr", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_block(name, S, res, far, frr):
    r = eo.eer_sweep(S)
    assert [r["EER"], r["thres"], r["FAR"], r["FRR"]] == list(res)
    assert np.array_equal(np.asarray(r["far"]), far) and np.array_equal(np.asarray(r["frr"]), frr)


def test_goldens_are_not_all_degenerate():
    """The reference keeps a threshold only when |FAR - FRR| < 1 with its own (large) denominators, so
    its final answer is usually 0 or the first clean threshold; the per-threshold curves are what pin
    the counting."""
    nz = sum(int(np.count_nonzero(c[3]) > 0 and np.count_nonzero(c[4]) > 0) for c in CASES)
    assert nz >= 5
    assert len({float(c[2][1]) for c in CASES}) >= 4        # several different selected thresholds


def _mixed_matrix(N, M, seed):
    """Cosine-like values all over [-0.2, 1.0] with exact hits on thresholds, a NaN and infinities."""
    rng = np.random.default_rng(seed)
    S = rng.uniform(-0.2, 1.0, size=(N, M, N)).astype(np.float32)
    idx = np.arange(N)
    S[idx, :, idx] = rng.uniform(0.4, 1.0, size=(N, M)).astype(np.float32)
    th = [np.float32(t) for t in eo.default_thresholds()]
    flat = S.reshape(-1)
    for q in range(0, min(flat.size, 200), 3):
        flat[q] = th[q % 50]                                 # exactly on a threshold: not accepted (strict >)
    if flat.size > 10:
        flat[5], flat[7], flat[9] = np.nan, np.inf, -np.inf
    return S


@pytest.mark.gpu
@pytest.mark.parametrize("name,S,res,far,frr", CASES, ids=[c[0] for c in CASES])
def test_gpu_sweep_matches_reference_golden(name, S, res, far, frr):
    import speaker_embedding_ge2e_loss_b200 as pkg
    r = pkg.eer_sweep(torch.tensor(S, device="cuda:0"))
    assert [r.EER, r.thres, r.FAR, r.FRR] == list(res)
    assert np.array_equal(np.asarray(r.far), far) and np.array_equal(np.asarray(r.frr), frr)


@pytest.mark.gpu
@pytest.mark.parametrize("N,M", [(2, 2), (4, 5), (7, 3), (33, 6), (64, 10), (100, 7), (257, 4), (1024, 10)])
def test_gpu_counts_bit_exact_vs_oracle(N, M):
    import speaker_embedding_ge2e_loss_b200 as pkg
    S = _mixed_matrix(N, M, seed=N + M)
    ths = eo.default_thresholds()
    a, o = pkg.threshold_counts(torch.tensor(S, device="cuda:0"), ths)
    ra, ro = eo.threshold_counts(S, ths)
    assert np.array_equal(a, ra) and np.array_equal(o, ro)
    r, rr = pkg.eer_sweep(torch.tensor(S, device="cuda:0")), eo.eer_sweep(S)
    assert (r.EER, r.thres, r.FAR, r.FRR) == (rr["EER"], rr["thres"], rr["FAR"], rr["FRR"])
    assert r.far == rr["far"] and r.frr == rr["frr"]


@pytest.mark.gpu
def test_gpu_unsorted_duplicate_and_single_thresholds():
    import speaker_embedding_ge2e_loss_b200 as pkg
    S = _mixed_matrix(48, 5, seed=3)
    Sd = torch.tensor(S, device="cuda:0")
    for ths in ([0.9, 0.1, 0.5, 0.5, -1.0, 2.0, 0.73], [0.61], list(np.linspace(-0.5, 1.5, 1024))):
        a, o = pkg.threshold_counts(Sd, ths)
        ra, ro = eo.threshold_counts(S, ths)
        assert np.array_equal(a, ra) and np.array_equal(o, ro)
    with pytest.raises(ValueError):
        pkg.threshold_counts(Sd[:, :, :5], [0.5])
    with pytest.raises(ValueError):
        pkg.eer_sweep(Sd[:1, :, :1])
    with pytest.raises(Exception):
        pkg.threshold_counts(Sd, list(np.linspace(0, 1, 1025)))       # more than 1024 thresholds


@pytest.mark.gpu
def test_gpu_evaluate_eer_is_s5_call_sequence():
    """evaluate_eer(E) == the script's sequence (s5:42-46) on the oracle's float32 similarity matrix;
    counts compared away from threshold ties (the matrix itself is fp32 arithmetic on both sides)."""
    import speaker_embedding_ge2e_loss_b200 as pkg
    N, M, D = 12, 6, 64
    E = orc.make_embeddings(N, M, D, seed=11, kind="clustered")
    res = pkg.evaluate_eer(torch.tensor(np.asarray(E, dtype=np.float32), device="cuda:0"))
    S = pkg.GE2ELoss.get_cos_sim(torch.tensor(np.asarray(E, dtype=np.float32), device="cuda:0"), None).cpu().numpy()
    ref = eo.eer_sweep(S)
    assert np.array_equal(res.accept_all, ref["accept_all"]) and np.array_equal(res.accept_own, ref["accept_own"])
    assert (res.EER, res.thres, res.FAR, res.FRR) == (ref["EER"], ref["thres"], ref["FAR"], ref["FRR"])
