"""Measurement for the EER sweep row: the count kernel over a device-resident
similarity matrix against the HBM roofline (it reads the matrix once: 4 bytes per entry), beside the
reference's own numpy sweep (s5_eval_model.py:57-89 restated in oracle/eer_oracle.py) on the host,
which also needs the matrix copied to the host first (s5:46)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speaker_embedding_ge2e_loss_b200 as pkg  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import _lib  # noqa: E402
from oracle import eer_oracle as eo  # noqa: E402

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs") or 7700.0
out = []
for (N, M) in [(64, 10), (1024, 10), (2048, 16)]:
    g = torch.Generator(device="cpu").manual_seed(N)
    n_rot = max(2, int(np.ceil(1.5 * 126e6 / (N * M * N * 4))))
    n_rot = min(n_rot, 64)
    mats = [(torch.rand((N, M, N), generator=g) * 1.2 - 0.2).to(dev) for _ in range(n_rot)]
    T = 50
    th = torch.tensor(np.asarray([np.float32(t) for t in eo.default_thresholds()], dtype=np.float32), device=dev)
    counts = torch.empty((2, T), dtype=torch.int64, device=dev)
    h = _lib.lib()
    nb = h.ge2e_b200_threshold_counts_scratch_bytes(T)
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run(k):
        global st
        h.ge2e_b200_threshold_counts(mats[k % n_rot].data_ptr(), N, M, th.data_ptr(), T, counts[0].data_ptr(),
                                     counts[1].data_ptr(), scratch.data_ptr(), nb, st)
    for k in range(5):
        run(k)
    torch.cuda.synchronize()
    # 20 calls over rotating matrices captured in one CUDA graph: the eager loop is host-bound (~20 us per
    # ctypes call + memset + launch), which would be charged to the kernel
    steps = 20
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph, stream=side):
        st_keep = st
        st = torch.cuda.current_stream().cuda_stream
        for k in range(steps):
            run(k)
        st = st_keep
    for _ in range(3):
        gph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * steps) * 1e3
    nbytes = N * M * N * 4
    # host side: the reference's sequence = D2H of the matrix + numpy sweep
    t0 = time.perf_counter()
    S = mats[0].cpu().numpy()
    t1 = time.perf_counter()
    ref = eo.eer_sweep(S)
    t2 = time.perf_counter()
    t3 = time.perf_counter()
    res = pkg.eer_sweep(mats[0])
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    assert np.array_equal(res.accept_all, ref["accept_all"]) and np.array_equal(res.accept_own, ref["accept_own"])
    out.append({"N": N, "M": M, "matrix_MB": nbytes / 1e6, "rotating_matrices": n_rot, "kernel_us": us,
                "achieved_GBps": nbytes / us / 1e3, "hbm_peak_GBps": hbm, "frac": nbytes / us / 1e3 / hbm,
                "public_api_ms": (t4 - t3) * 1e3, "cpu_d2h_ms": (t1 - t0) * 1e3, "cpu_numpy_sweep_ms": (t2 - t1) * 1e3,
                "cpu_cores": os.cpu_count()})
    print(json.dumps(out[-1]))
    del mats
    torch.cuda.empty_cache()
