"""CPU tests: the oracle (numpy fp64) and the torch-CPU port are pinned against vectors
produced by the real reference class (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_names
from oracle import ge2e_oracle as orc
from oracle import ge2e_ref_port as port

NAMES = golden_names()


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_kat_values_match_survey(golden):
    # values derived from the reference on the s3:144-145 input
    k = golden["kat_w1_b0"]
    assert abs(k.r32["loss"] - 5.250094413757324) < 1e-6
    np.testing.assert_allclose(k.r32["per"].ravel(),
                               [1.5514450, 1.0986127, 0.7485732, 0.7485732, 0.5514450, 0.5514450],
                               rtol=1e-6)
    k = golden["kat_w10_bm5"]
    assert abs(k.r32["loss"] - 11.203167915344238) < 1e-5
    assert abs(k.r32["dw"] - 0.9699187875) < 1e-6


@pytest.mark.parametrize("name", NAMES)
def test_oracle_softmax_matches_reference_fp64(golden, name):
    c = golden[name]
    r = orc.forward_backward(c.E, c.w, c.b, 1e-6, orc.SOFTMAX)
    ref = c.r64
    assert abs(r["loss"] - ref["loss"]) <= 1e-11 * max(1.0, abs(ref["loss"]))
    np.testing.assert_allclose(r["per"], ref["per"], rtol=1e-10, atol=1e-12)
    assert rel(r["dE"], ref["dE"]) < 2e-7          # fixture dE stored as fp32
    assert abs(r["dw"] - ref["dw"]) <= 1e-9 * max(1.0, abs(ref["dw"]))
    assert abs(r["db"] - ref["db"]) <= 1e-9 * c.E.shape[0] * c.E.shape[1]
    # forward-only entry point agrees with the fwd+bwd one
    loss, per = orc.forward(c.E, c.w, c.b, 1e-6, orc.SOFTMAX)
    assert abs(loss - r["loss"]) <= 1e-12 * max(1.0, abs(loss))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_softmax_close_to_reference_fp32(golden, name):
    # what a user of the reference actually sees (fp32 eager); tolerance = fp32 noise
    c = golden[name]
    r = orc.forward_backward(c.E, c.w, c.b, 1e-6, orc.SOFTMAX)
    ref = c.r32
    assert abs(r["loss"] - ref["loss"]) <= 2e-5 * max(1.0, abs(ref["loss"]))
    # N == 1 is purely eps-driven (loss = log(1 + eps*exp(-S))): the reference's own fp32
    # autograd is only good to ~1e-4 there.
    tol = 2e-4 if name.startswith("onespk") else 2e-5
    assert rel(r["dE"], ref["dE"]) < tol
    assert abs(r["dw"] - ref["dw"]) <= 2e-5 * max(1.0, abs(ref["dw"]))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_contrast_matches_torch_autograd(golden, name):
    c = golden[name]
    r = orc.forward_backward(c.E, c.w, c.b, 1e-6, orc.CONTRAST)
    ref = c.rc
    assert abs(r["loss"] - ref["loss"]) <= 1e-11 * max(1.0, abs(ref["loss"]))
    assert rel(r["dE"], ref["dE"]) < 2e-7
    assert abs(r["dw"] - ref["dw"]) <= 1e-9 * max(1.0, abs(ref["dw"]))
    assert abs(r["db"] - ref["db"]) <= 1e-9 * max(1.0, abs(ref["db"]))


@pytest.mark.parametrize("name", [n for n in NAMES if "64x10" not in n and "n130" not in n] +
                         ["cfg2_64x10x256_random"])
def test_ref_port_matches_reference_fp32(golden, name):
    c = golden[name]
    torch.set_num_threads(1)
    E = torch.tensor(c.E, requires_grad=True)
    w = torch.tensor(c.w, requires_grad=True)
    b = torch.tensor(c.b, requires_grad=True)
    loss = port.loss_full(E, w, b, 1e-6)
    loss.backward()
    ref = c.r32
    assert abs(loss.item() - ref["loss"]) <= 1e-6 * max(1.0, abs(ref["loss"]))
    assert rel(E.grad.numpy(), ref["dE"]) < 1e-6
    assert abs(w.grad.item() - ref["dw"]) <= 1e-5 * max(1.0, abs(ref["dw"]))


def test_ref_port_row_sample_is_partial_sum(golden):
    c = golden["cfg1_4x8x256_random"]
    E = torch.tensor(c.E)
    rows = [0, 5, 9, 17, 31]
    part = port.loss_row_sample(E, torch.tensor(c.w), torch.tensor(c.b), rows, 1e-6).item()
    want = c.r64["per"].ravel()[rows].sum()
    assert abs(part - want) < 1e-4


def test_static_helpers_and_edge_cases():
    E = orc.make_embeddings(3, 4, 8, seed=11, kind="raw")
    np.testing.assert_allclose(orc.get_centroids(E), E.astype(np.float64).mean(1))
    u = orc.get_utterance_centroids(E)
    np.testing.assert_allclose(u[1, 2], (E[1].astype(np.float64).sum(0) - E[1, 2]) / 3, rtol=1e-12)
    # all-zero embeddings are finite (cos = 0): SURVEY 8(a-bis) item 13
    r = orc.forward_backward(np.zeros((2, 3, 4), np.float32))
    assert np.isfinite(r["loss"]) and np.isfinite(r["dE"]).all()
    # M == 1 is NaN in the reference (division by zero at s3:110-111)
    with np.errstate(all="ignore"):
        loss, _ = orc.forward(np.ones((2, 1, 4), np.float32))
    assert np.isnan(loss)


def test_oracle_gradient_is_consistent_with_finite_differences():
    E = orc.make_embeddings(3, 3, 6, seed=3, kind="raw").astype(np.float64)
    for variant in (orc.SOFTMAX, orc.CONTRAST):
        r = orc.forward_backward(E, 4.0, -1.0, 1e-6, variant, g=0.7)
        h = 1e-6
        num = np.zeros_like(E)
        for idx in np.ndindex(*E.shape):
            Ep, Em = E.copy(), E.copy()
            Ep[idx] += h
            Em[idx] -= h
            num[idx] = 0.7 * (orc.forward(Ep, 4.0, -1.0, 1e-6, variant)[0] -
                              orc.forward(Em, 4.0, -1.0, 1e-6, variant)[0]) / (2 * h)
        assert rel(r["dE"], num) < 1e-6
        dw = 0.7 * (orc.forward(E, 4.0 + h, -1.0, 1e-6, variant)[0] -
                    orc.forward(E, 4.0 - h, -1.0, 1e-6, variant)[0]) / (2 * h)
        db = 0.7 * (orc.forward(E, 4.0, -1.0 + h, 1e-6, variant)[0] -
                    orc.forward(E, 4.0, -1.0 - h, 1e-6, variant)[0]) / (2 * h)
        assert abs(dw - r["dw"]) < 1e-6 and abs(db - r["db"]) < 1e-6


# ------------------------------------------------------------------ chunked torch restatement (large-N checker)
@pytest.mark.parametrize("N,M,D,kind,variant,wbg", [
    (64, 10, 64, "clustered", "softmax", (10.0, -5.0, 1.0)), (33, 3, 48, "raw", "softmax", (-3.0, 0.5, 0.25)),
    (40, 5, 32, "random", "contrast", (10.0, -5.0, 1.0)), (7, 2, 16, "clustered", "contrast", (4.0, -1.0, -0.7)),
])
def test_torch_oracle_matches_numpy_oracle(N, M, D, kind, variant, wbg):
    """oracle/ge2e_oracle_torch.py (what the GPU tests evaluate in float64 at configs 3 / 4) is the numpy
    oracle row chunk by row chunk: identical to 1e-12 whatever the chunk size."""
    import torch
    from oracle import ge2e_oracle_torch as orct
    w, b, g = wbg
    E = orc.make_embeddings(N, M, D, seed=N + D, kind=kind)
    ref = orc.forward_backward(E, w, b, 1e-6, variant, g=g)
    for chunk in (N * M, 37):
        got = orct.forward_backward(torch.tensor(E), w, b, 1e-6, variant, g=g, chunk=chunk)
        assert abs(got["loss"] - ref["loss"]) <= 1e-12 * max(1.0, abs(ref["loss"]))
        assert np.linalg.norm(got["dE"].numpy() - ref["dE"]) <= 1e-12 * np.linalg.norm(ref["dE"])
        assert abs(got["dw"] - ref["dw"]) <= 1e-12 * max(1.0, abs(ref["dw"]))
        assert abs(got["db"] - ref["db"]) <= 1e-12 * max(1.0, abs(ref["db"]))
        assert np.allclose(got["per"].numpy(), ref["per"], rtol=1e-12, atol=1e-14)
