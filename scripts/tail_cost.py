"""Measurement for the model-tail row: the fused Linear + L2-normalise kernel
against its rooflines (tensor: 2 U H D flops at the measured TF32 rate; HBM: X + W + E bytes once) and
beside the reference's three torch lines (s2:30-34) on the same GPU, forward and forward+backward."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speaker_embedding_ge2e_loss_b200 as pkg  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
hbm, tf32 = pk["hbm_gbs"], pk["bf16_tflops"] / 2
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, steps=30, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


for (U, frames, H, D) in [(640, 160, 768, 256), (10240, 1, 768, 256), (131072, 1, 768, 256)]:
    out = torch.randn(U, frames, H, device=dev)
    lin = torch.nn.Linear(H, D).to(dev)
    W, b = lin.weight.detach(), lin.bias.detach()
    x = out[:, frames - 1]
    E = torch.empty(U, D, device=dev)
    inv = torch.empty(U, device=dev)
    h = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream

    def ours_fwd():
        h.ge2e_b200_embed_tail_fwd(x.data_ptr(), x.stride(0), W.data_ptr(), b.data_ptr(), U, H, D, E.data_ptr(),
                                   inv.data_ptr(), st)

    def ref_fwd():
        y = torch.nn.functional.linear(out[:, out.size(1) - 1], W, b)
        return y / torch.norm(y, dim=1).unsqueeze(1)

    res = {"U": U, "frames": frames, "H": H, "D": D}
    res["ours_fwd_us"] = timed(ours_fwd)
    torch.backends.cuda.matmul.allow_tf32 = False
    res["torch_fp32_fwd_us"] = timed(ref_fwd)
    torch.backends.cuda.matmul.allow_tf32 = True
    res["torch_tf32_fwd_us"] = timed(ref_fwd)
    flops = 2.0 * U * H * D
    nbytes = 4.0 * (U * H + D * H + U * D)
    t = res["ours_fwd_us"] * 1e-6
    res["roofline"] = {"tensor_TFLOPs": flops / t / 1e12, "tensor_peak": tf32, "tensor_frac": flops / t / 1e12 / tf32,
                       "hbm_GBps": nbytes / t / 1e9, "hbm_peak": hbm, "hbm_frac": nbytes / t / 1e9 / hbm,
                       "bound": "hbm" if nbytes / hbm / 1e9 > flops / tf32 / 1e12 else "tensor"}
    # the two gradient GEMMs of the Linear layer (dX = dY W, dW = dY^T X): one tcgen05 kernel each vs library GEMMs
    dY = torch.randn(U, D, device=dev)
    dX = torch.empty(U, H, device=dev)
    dW = torch.empty(D, H, device=dev)

    def ours_gemms():
        rc = h.ge2e_b200_embed_tail_bwd_gemms(dY.data_ptr(), W.data_ptr(), x.data_ptr(), x.stride(0), U, H, D,
                                              dX.data_ptr(), H, dW.data_ptr(), st)
        assert rc == 0, rc

    def lib_gemms():
        torch.matmul(dY, W, out=dX)
        torch.matmul(dY.t(), x, out=dW)

    res["ours_bwd_gemms_us"] = timed(ours_gemms)
    torch.backends.cuda.matmul.allow_tf32 = False
    res["torch_fp32_bwd_gemms_us"] = timed(lib_gemms)
    torch.backends.cuda.matmul.allow_tf32 = True
    res["torch_tf32_bwd_gemms_us"] = timed(lib_gemms)
    gf = 4.0 * U * H * D
    gb = 4.0 * (2 * U * D + 2 * D * H + 2 * U * H)
    tg = res["ours_bwd_gemms_us"] * 1e-6
    res["bwd_gemms_roofline"] = {"tensor_frac": gf / tg / 1e12 / tf32, "hbm_frac": gb / tg / 1e9 / hbm,
                                 "bound": "hbm" if gb / hbm / 1e9 > gf / tf32 / 1e12 else "tensor"}
    # forward + backward through the public module vs the same three lines under autograd
    tail = pkg.ProjectionL2Norm(H, D).to(dev)
    dE = torch.randn(U, D, device=dev)
    o1 = out.clone().requires_grad_(True)

    def ours_fb():
        o1.grad = None
        tail.projection.weight.grad = tail.projection.bias.grad = None
        (tail(o1) * dE).sum().backward()

    def ref_fb():
        o1.grad = None
        tail.projection.weight.grad = tail.projection.bias.grad = None
        y = tail.projection(o1[:, o1.size(1) - 1])
        ((y / torch.norm(y, dim=1).unsqueeze(1)) * dE).sum().backward()

    if U * frames * H * 4 < 8e9:
        torch.backends.cuda.matmul.allow_tf32 = False
        res["ours_fwd_bwd_us"] = timed(ours_fb, steps=15)
        res["torch_fp32_fwd_bwd_us"] = timed(ref_fb, steps=15)
    print(json.dumps(res))
    del out, o1
    torch.cuda.empty_cache()
