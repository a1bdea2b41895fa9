"""Model tail (s2_model_GE2E_loss_speach_embed.py:28-34): last frame -> Linear ->
L2 normalise.

CPU: the fp64 oracle against vectors produced by the reference's real model class
(tests/golden/make_tail_golden.py).  GPU: the tcgen05 kernel (TF32 operands, fp32 accumulation) and
its backward against the oracle, tolerance 2e-3 relative (the TF32 tolerance of the loss path;
observed ~3e-4) on E and every gradient.
"""
import os

import numpy as np
import pytest
import torch

from oracle import tail_oracle as to

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tail_reference_vectors.npz")
TF32_TOL = 2e-3


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def golden_cases():
    z = np.load(GOLD)
    keys = sorted({k[:-len("_x_last")] for k in z.files if k.endswith("_x_last")})
    return [(k, {f: z[f"{k}_{f}"] for f in ("x_last", "W", "bias", "E", "dE", "dX", "dW", "dbias")}) for k in keys]


CASES = golden_cases()


@pytest.mark.parametrize("name,g", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_model(name, g):
    tol = 1e-12 if name.endswith("f64") else 5e-6
    e, _ = to.embed_tail(g["x_last"], g["W"], g["bias"])
    b = to.embed_tail_backward(g["x_last"], g["W"], g["bias"], g["dE"])
    assert rel(e, g["E"]) <= tol
    assert rel(b["dX"], g["dX"]) <= tol and rel(b["dW"], g["dW"]) <= tol and rel(b["dbias"], g["dbias"]) <= tol
    assert np.allclose(np.linalg.norm(e, axis=1), 1.0, atol=1e-12)


def _run_cuda(x_last, W, bias, dE, frames=None):
    import speaker_embedding_ge2e_loss_b200 as pkg
    dev = torch.device("cuda:0")
    Wt = torch.tensor(np.asarray(W, dtype=np.float32), device=dev, requires_grad=True)
    bt = None if bias is None else torch.tensor(np.asarray(bias, dtype=np.float32), device=dev, requires_grad=True)
    if frames is None:
        x = torch.tensor(np.asarray(x_last, dtype=np.float32), device=dev, requires_grad=True)
    else:                       # LSTM-output shaped input: the kernel reads the last frame through the row stride
        full = np.random.default_rng(0).standard_normal((x_last.shape[0], frames, x_last.shape[1])).astype(np.float32)
        full[:, -1] = x_last
        x = torch.tensor(full, device=dev, requires_grad=True)
    E = pkg.project_normalize(x, Wt, bt)
    (E * torch.tensor(np.asarray(dE, dtype=np.float32), device=dev)).sum().backward()
    torch.cuda.synchronize()
    dX = x.grad if frames is None else x.grad[:, -1]
    if frames is not None:
        assert float(x.grad[:, :-1].abs().max()) == 0.0
    return dict(E=E.detach().cpu().numpy(), dX=dX.cpu().numpy(), dW=Wt.grad.cpu().numpy(),
                dbias=None if bt is None else bt.grad.cpu().numpy())


def _check(got, x_last, W, bias, dE):
    e, _ = to.embed_tail(x_last, W, bias)
    b = to.embed_tail_backward(x_last, W, bias, dE)
    assert rel(got["E"], e) <= TF32_TOL, rel(got["E"], e)
    assert rel(got["dX"], b["dX"]) <= TF32_TOL, rel(got["dX"], b["dX"])
    assert rel(got["dW"], b["dW"]) <= TF32_TOL, rel(got["dW"], b["dW"])
    if bias is not None:
        assert rel(got["dbias"], b["dbias"]) <= TF32_TOL, rel(got["dbias"], b["dbias"])
    assert np.allclose(np.linalg.norm(got["E"], axis=1), 1.0, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name,g", [c for c in CASES if c[0].endswith("f32")], ids=[c[0] for c in CASES if c[0].endswith("f32")])
def test_gpu_matches_reference_golden(name, g):
    got = _run_cuda(g["x_last"], g["W"], g["bias"], g["dE"])
    assert rel(got["E"], g["E"]) <= TF32_TOL and rel(got["dX"], g["dX"]) <= TF32_TOL
    assert rel(got["dW"], g["dW"]) <= TF32_TOL and rel(got["dbias"], g["dbias"]) <= TF32_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("U,H,D,frames,use_bias", [
    (1, 32, 64, None, True), (127, 768, 256, None, True), (128, 768, 256, 3, True), (129, 40, 128, None, False),
    (640, 768, 256, 5, True), (1000, 100, 256, None, True), (10240, 768, 256, None, True), (300, 36, 64, 2, False)])
def test_gpu_vs_oracle_seeded(U, H, D, frames, use_bias):
    rng = np.random.default_rng(U + H + D)
    x = rng.standard_normal((U, H)).astype(np.float32)
    W = (rng.standard_normal((D, H)) / np.sqrt(H)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(D)).astype(np.float32) if use_bias else None
    dE = rng.standard_normal((U, D)).astype(np.float32)
    _check(_run_cuda(x, W, bias, dE, frames), x, W, bias, dE)


@pytest.mark.gpu
def test_gpu_module_matches_reference_lines_and_loads_state_dict():
    """ProjectionL2Norm against the reference's three lines written with torch ops on the same weights;
    state-dict keys are the reference model's (projection.weight / projection.bias)."""
    import speaker_embedding_ge2e_loss_b200 as pkg
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    tail = pkg.ProjectionL2Norm(768, 256).to(dev)
    assert sorted(tail.state_dict().keys()) == ["projection.bias", "projection.weight"]
    out = torch.randn(64, 9, 768, device=dev)
    E = tail(out)
    x = out[:, out.size(1) - 1]
    y = torch.nn.functional.linear(x.double(), tail.projection.weight.double(), tail.projection.bias.double())
    ref = y / torch.norm(y, dim=1).unsqueeze(1)
    assert rel(E.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= TF32_TOL
    # feeds the loss unchanged
    crit = pkg.GE2ELoss(None, device=dev, precision="tf32")
    loss = crit(E.view(8, 8, 256))
    loss.backward()
    assert torch.isfinite(loss) and tail.projection.weight.grad is not None
    assert float(tail.projection.weight.grad.abs().max()) > 0


@pytest.mark.gpu
def test_gpu_backward_gemms_run_on_our_kernels():
    """dX = dY W and dW = dY^T X are launches of this library (tcgen05 kernel), not library GEMMs, wherever the
    hidden size is a multiple of 32 (the reference's 768): forward 1 launch, backward 3 (rows, dX, dW)."""
    import speaker_embedding_ge2e_loss_b200 as pkg
    dev = torch.device("cuda:0")
    h = pkg.lib()
    assert h.ge2e_b200_embed_tail_bwd_gemms_supported(10240, 768, 256, 768, 768) == 1
    assert h.ge2e_b200_embed_tail_bwd_gemms_supported(640, 768, 256, 160 * 768, 160 * 768) == 1     # last-frame strides
    assert h.ge2e_b200_embed_tail_bwd_gemms_supported(100, 36, 64, 36, 36) == 0
    tail = pkg.ProjectionL2Norm(768, 256).to(dev)
    out = torch.randn(300, 3, 768, device=dev, requires_grad=True)
    n0 = h.ge2e_b200_launch_count()
    E = tail(out)
    n1 = h.ge2e_b200_launch_count()
    E.sum().backward()
    torch.cuda.synchronize()
    n2 = h.ge2e_b200_launch_count()
    assert n1 - n0 == 1 and n2 - n1 == 3, (n1 - n0, n2 - n1)
    assert float(out.grad[:, :-1].abs().max()) == 0.0 and float(out.grad[:, -1].abs().max()) > 0


@pytest.mark.gpu
def test_gpu_unsupported_shapes_raise():
    import speaker_embedding_ge2e_loss_b200 as pkg
    dev = torch.device("cuda:0")
    with pytest.raises(RuntimeError, match="not supported"):
        pkg.project_normalize(torch.randn(8, 64, device=dev), torch.randn(100, 64, device=dev))     # D not 64/128/256
    with pytest.raises(RuntimeError, match="not supported"):
        pkg.project_normalize(torch.randn(8, 30, device=dev), torch.randn(64, 30, device=dev))      # H % 4 != 0
    with pytest.raises(RuntimeError):
        pkg.project_normalize(torch.randn(8, 64), torch.randn(64, 64))                              # CPU tensors
