"""One trainer iteration (s4_train_embed_model.py:170-203) with every widened row in place: batch from the
device-resident bank (row 4), LSTM (torch), fused projection + L2 normalise (row 2), loss with the fused
unperm gather, clip + SGD tail for w / b (row 1) -- against the same iteration written with the
reference's torch lines around the fp64 oracle loss."""
import random

import numpy as np
import pytest
import torch

from oracle import batch_oracle as bo
from oracle import ge2e_oracle as orc

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_composed_training_iteration_matches_reference_sequence():
    import speaker_embedding_ge2e_loss_b200 as pkg
    dev = torch.device("cuda:0")
    S, frames, mels, N, M, L, H, D = 6, 40, 8, 4, 5, 24, 64, 64
    rng = np.random.default_rng(5)
    files = [rng.standard_normal((int(rng.integers(3, 7)), frames, mels)) for _ in range(S)]
    bank = pkg.SpectrogramBank(files, device=dev)
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(mels, H, num_layers=2, batch_first=True).to(dev)
    tail = pkg.ProjectionL2Norm(H, D).to(dev)
    crit = pkg.GE2ELoss(None, device=dev, precision="fp32")
    speakers = [4, 1, 5, 0]

    np.random.seed(3)
    random.seed(3)
    batch, unperm = bank.training_batch(speakers, M, L)
    out, _ = lstm(batch)
    E = tail(out)
    loss = crit(E, unperm=unperm, speakers=N)
    loss.backward()
    g_ours = [p.grad.clone() for p in lstm.parameters()]
    gW_ours = tail.projection.weight.grad.clone()
    w0, b0, dw, db = crit.w.item(), crit.b.item(), crit.w.grad.item(), crit.b.grad.item()
    crit.clip_and_sgd_step(lr=0.01, max_norm=1.0)
    torch.cuda.synchronize()

    # the reference's sequence: dataset crop + collate + reshape + perm + float, model lines, unperm, fp64 loss
    np.random.seed(3)
    random.seed(3)
    spk_files = [files[s] for s in speakers]
    utt, clip = bo.draw_indices([a.shape[0] for a in spk_files], M, frames, L)
    perm = random.sample(range(0, N * M), N * M)
    ref_batch = torch.tensor(bo.assemble(spk_files, utt, clip, L, perm), device=dev)
    assert torch.equal(batch, ref_batch)
    for p in list(lstm.parameters()) + list(tail.parameters()):
        p.grad = None
    out2, _ = lstm(ref_batch)
    y = tail.projection(out2[:, out2.size(1) - 1])
    e = y / torch.norm(y, dim=1).unsqueeze(1)
    un = [0] * (N * M)
    for i, j in enumerate(perm):
        un[j] = i
    e_nm = e[un].reshape(N, M, D)
    ref = orc.forward_backward(e_nm.detach().cpu().numpy(), 10.0, -5.0, 1e-6, "softmax")
    e_nm.backward(torch.tensor(ref["dE"], dtype=torch.float32, device=dev))
    torch.cuda.synchronize()

    assert abs(loss.item() - ref["loss"]) <= 2e-3 * abs(ref["loss"])
    assert abs(dw - ref["dw"]) <= 2e-3 * max(1.0, abs(ref["dw"])) and abs(db - ref["db"]) <= 2e-3 * max(1.0, abs(ref["db"]))
    assert rel(gW_ours.cpu().numpy(), tail.projection.weight.grad.cpu().numpy()) <= 5e-3
    for a, p in zip(g_ours, lstm.parameters()):
        assert rel(a.cpu().numpy(), p.grad.cpu().numpy()) <= 5e-3
    coef = min(1.0, 1.0 / (np.hypot(dw, db) + 1e-6))
    assert abs(crit.w.item() - (w0 - 0.01 * dw * coef)) <= 1e-5 and abs(crit.b.item() - (b0 - 0.01 * db * coef)) <= 1e-5
