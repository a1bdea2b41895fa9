"""BASELINE config 5 with the data path included: one whole trainer iteration (s4_train_embed_model.py:
168-205) at N=64 speakers x M=10 utterances, 160 x 40 crops, 3-layer LSTM 768 + Linear 256 --
  reference sequence: np.load per speaker + crop (s1:59-77), collate, float64 H2D, reshape, perm, .float(),
                      LSTM, torch projection + normalise, [unperm], eager PyTorch GE2E in the reference's
                      formulation, backward, two clip_grad_norm_, SGD step, loss to the host;
  this repo:          SpectrogramBank.training_batch (row 4), LSTM, ProjectionL2Norm (row 2), GE2ELoss with
                      the fused unperm gather, backward, clip of the model, clip + SGD of w / b in one kernel
                      (row 1), SGD step of the model, loss to the host.
Wall clock per iteration (the data path is host work), synthetic spectrogram files in a temp directory.
    python scripts/train_iteration.py [N M]"""
import json
import os
import random
import sys
import tempfile
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speaker_embedding_ge2e_loss_b200 as pkg  # noqa: E402
sys.argv = sys.argv[:1] + [str(a) for a in sys.argv[1:3]]
from train_step_share import EagerGE2E  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L, FRAMES, MEL, HID, EMB, LAYERS, UTTS = 160, 180, 40, 768, 256, 3, 12
dev = torch.device("cuda:0")


def reference_iteration(files, lstm, proj, crit, opt):
    items = []
    for f in files:                                                     # Dataset.__getitem__ x N (s1:59-77)
        spr = np.load(f)
        idx = np.random.randint(0, spr.shape[0], M)
        clip = np.random.randint(0, spr.shape[1] - L - 1)
        items.append(spr[idx, :, :][:, clip:clip + L, :])
    batch = torch.tensor(np.stack(items)).to(dev)                       # collate + s4:170
    batch = torch.reshape(batch, (N * M, L, MEL))
    perm = random.sample(range(0, N * M), N * M)
    unperm = list(perm)
    for i, j in enumerate(perm):
        unperm[j] = i
    x, _ = lstm(batch[perm].float())                                    # s2:28
    y = proj(x[:, x.size(1) - 1])
    emb = (y / torch.norm(y, dim=1).unsqueeze(1))[unperm]
    loss = crit(torch.reshape(emb, (N, M, EMB)))
    opt.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(list(lstm.parameters()) + list(proj.parameters()), 3.0)
    torch.nn.utils.clip_grad_norm_(crit.parameters(), 1.0)
    opt.step()
    return float(loss.to("cpu").detach().numpy())


def our_iteration(bank, speakers, lstm, tail, crit, opt):
    batch, unperm = bank.training_batch(speakers, M, L)
    x, _ = lstm(batch)
    emb = tail(x)
    loss = crit(emb, unperm=unperm, speakers=N)
    opt.zero_grad()
    crit.w.grad = crit.b.grad = None
    loss.backward()
    torch.nn.utils.clip_grad_norm_(list(lstm.parameters()) + list(tail.parameters()), 3.0)
    crit.clip_and_sgd_step(lr=0.01, max_norm=1.0)
    opt.step()
    return float(loss.to("cpu").detach().numpy())


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


with tempfile.TemporaryDirectory() as d:
    rng = np.random.default_rng(0)
    files = []
    for s in range(N):
        p = os.path.join(d, f"sv_{s:04d}.npy")
        np.save(p, rng.standard_normal((UTTS, FRAMES, MEL)))            # float64, as s0 writes it
        files.append(p)
    torch.manual_seed(0)
    lstm = nn.LSTM(MEL, HID, num_layers=LAYERS, batch_first=True).to(dev)
    proj = nn.Linear(HID, EMB).to(dev)
    ref_crit = EagerGE2E()
    opt = torch.optim.SGD([{"params": list(lstm.parameters()) + list(proj.parameters())}, {"params": ref_crit.parameters()}], lr=0.01)
    t_ref = timed(lambda: reference_iteration(files, lstm, proj, ref_crit, opt))

    t0 = time.perf_counter()
    bank = pkg.SpectrogramBank.from_dir(d, device=dev)
    t_bank = (time.perf_counter() - t0) * 1e3
    tail = pkg.ProjectionL2Norm(HID, EMB).to(dev)
    crit = pkg.GE2ELoss(None, device=dev, precision="tf32")
    opt2 = torch.optim.SGD(list(lstm.parameters()) + list(tail.parameters()), lr=0.01)
    speakers = list(range(N))
    t_ours = timed(lambda: our_iteration(bank, speakers, lstm, tail, crit, opt2))
    print(json.dumps({"N": N, "M": M, "crop": [L, MEL], "lstm": [MEL, HID, LAYERS], "emb": EMB,
                      "reference_sequence_ms_per_iteration": t_ref, "this_repo_ms_per_iteration": t_ours,
                      "speedup": t_ref / t_ours, "bank_build_ms_once": t_bank, "host_cores": os.cpu_count()}))
