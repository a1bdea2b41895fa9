"""Measurement for the batch-assembly row: the gather kernel against the HBM
roofline (reads + writes the batch once: 8 bytes per output float) and, beside it, the reference's
sequence on the host cores -- numpy crop per speaker (s1:59-77), collate, float64 H2D, reshape,
permute, .float() (s4:170-186, s2:28)."""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speaker_embedding_ge2e_loss_b200 as pkg  # noqa: E402
from oracle import batch_oracle as bo  # noqa: E402

dev = torch.device("cuda:0")
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (N, M, L, frames, mels, utts) in [(64, 10, 160, 180, 40, 8), (1024, 10, 160, 180, 40, 8)]:
    rng = np.random.default_rng(N)
    files = [rng.standard_normal((utts, frames, mels)) for _ in range(N)]
    bank = pkg.SpectrogramBank(files, device=dev)
    speakers = list(range(N))
    out = torch.empty((N * M, L, mels), device=dev)
    utt, clip = bank.draw(speakers, M, L)
    perm = random.sample(range(N * M), N * M)
    off = ((bank.first[np.asarray(speakers)][:, None] + utt) * frames + clip[:, None]) * mels
    off_dev = torch.from_numpy(off.reshape(-1)[np.asarray(perm)].copy()).to(dev)
    h = pkg.lib()
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for it in range(25):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.ge2e_b200_gather_spans(bank.data.data_ptr(), off_dev.data_ptr(), N * M, L * mels, 1, out.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        if it >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    nbytes = 2.0 * N * M * L * mels * 4
    t0 = time.perf_counter()
    b2, unperm = bank.training_batch(speakers, M, L)
    torch.cuda.synchronize()
    api_ms = (time.perf_counter() - t0) * 1e3
    # the reference's sequence on the host + its device tail
    t0 = time.perf_counter()
    u2, c2 = bo.draw_indices([utts] * N, M, frames, L)
    items = np.stack([bo.get_item(files[s], u2[s], int(c2[s]), L) for s in speakers])      # float64 [N, M, L, mels]
    t1 = time.perf_counter()
    d = torch.from_numpy(items).to(dev)
    d = torch.reshape(d, (N * M, L, mels))[perm].float()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(json.dumps({"N": N, "M": M, "crop": L, "mels": mels, "batch_MB": nbytes / 2e6, "kernel_us": us,
                      "achieved_GBps": nbytes / us / 1e3, "hbm_peak_GBps": hbm, "frac": nbytes / us / 1e3 / hbm,
                      "public_api_ms": api_ms, "ref_host_numpy_ms": (t1 - t0) * 1e3,
                      "ref_h2d_f64_reshape_perm_float_ms": (t2 - t1) * 1e3, "cpu_cores": os.cpu_count()}))
