"""The embedder's tail, fused for the loss.

Reference: embedding_model_GE2E/s2_model_GE2E_loss_speach_embed.py:28-34 -- after the LSTM stack the
model keeps the last frame, applies ``self.projection = nn.Linear(hidden, embedding)`` and divides by
the row norm.  ``ProjectionL2Norm`` holds the same ``projection`` sub-module (state-dict keys
``projection.weight`` / ``projection.bias``, so the reference's checkpoints load) and runs the three
lines as ONE tcgen05 kernel (``ge2e_b200_embed_tail_fwd``): the last-frame select is the row stride
of the TMA tensor map, the normalisation is the GEMM's epilogue, and the un-normalised projection
never reaches HBM.  TF32 tensor cores, fp32 accumulation.  CUDA (sm_100a) only; no fallback.  The backward is
the row Jacobian of the normalisation (``ge2e_b200_embed_tail_bwd_rows``) followed by the two gradient GEMMs of
the Linear layer, dX = dY W and dW = dY^T X, as one more tcgen05 kernel (``ge2e_b200_embed_tail_bwd_gemms``; hidden
sizes that are not a multiple of 32 take library GEMMs instead).

    tail = ProjectionL2Norm(768, 256).to("cuda")
    out, _ = lstm(mel)                      # [U, frames, 768]
    E = tail(out)                           # [U, 256], unit rows  ==  model.forward's return value
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops


def _rows_view(x: torch.Tensor) -> torch.Tensor:
    """[U, H] view the kernel can address: unit column stride, 16-byte aligned rows."""
    if x.dim() == 3:
        x = x[:, x.size(1) - 1]                      # s2:30, a strided view -- no copy
    if x.dim() != 2:
        raise ValueError(f"expected [U, H] or [U, frames, H], got {tuple(x.shape)}")
    if x.dtype != torch.float32:
        x = x.float()                                # s2:31 `.float()`
    if x.stride(1) != 1 or x.stride(0) % 4 != 0 or x.data_ptr() % 16 != 0 or x.stride(0) < x.size(1):
        x = x.contiguous()
    return x


class _ProjectNormalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ops._need_cuda(x, weight)
        xv = _rows_view(x.detach())
        U, H = xv.shape
        D = weight.shape[0]
        if weight.shape[1] != H:
            raise ValueError(f"weight must be [D, {H}], got {tuple(weight.shape)}")
        w = weight.detach().float().contiguous()
        b = None if bias is None else bias.detach().float().contiguous()
        E = torch.empty((U, D), dtype=torch.float32, device=xv.device)
        inv = torch.empty(U, dtype=torch.float32, device=xv.device)
        with torch.cuda.device(xv.device):
            _lib.check(_lib.lib().ge2e_b200_embed_tail_fwd(xv.data_ptr(), xv.stride(0), w.data_ptr(),
                                                           None if b is None else b.data_ptr(), U, H, D,
                                                           E.data_ptr(), inv.data_ptr(), ops._stream()),
                       "ge2e_b200_embed_tail_fwd")
        ctx.save_for_backward(xv, w, E, inv)
        ctx.has_bias = bias is not None
        ctx.x_shape = tuple(x.shape)
        return E

    @staticmethod
    def backward(ctx, dE):
        xv, w, E, inv = ctx.saved_tensors
        U, D = E.shape
        g = dE.float().contiguous()
        dY = torch.empty_like(E)
        dbias = torch.empty(D, dtype=torch.float32, device=E.device) if ctx.has_bias else None
        with torch.cuda.device(E.device):
            _lib.check(_lib.lib().ge2e_b200_embed_tail_bwd_rows(g.data_ptr(), E.data_ptr(), inv.data_ptr(), U, D,
                                                                dY.data_ptr(),
                                                                None if dbias is None else dbias.data_ptr(),
                                                                ops._stream()), "ge2e_b200_embed_tail_bwd_rows")
        dX = dW = None
        H = xv.shape[1]
        want_x, want_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if want_x:                                                  # gradient of the last-frame select: zeros elsewhere
            dX = (torch.zeros if len(ctx.x_shape) == 3 else torch.empty)(ctx.x_shape, dtype=torch.float32, device=E.device)
            dx_last = dX[:, ctx.x_shape[1] - 1] if len(ctx.x_shape) == 3 else dX
        if want_w:
            dW = torch.empty((D, H), dtype=torch.float32, device=E.device)
        h = _lib.lib()
        dxs = dx_last.stride(0) if want_x else H
        if (want_x or want_w) and h.ge2e_b200_embed_tail_bwd_gemms_supported(U, H, D, xv.stride(0), dxs) == 1:
            # dX = dY W and dW = dY^T X on tcgen05 (dX lands in the last frame through the row stride)
            with torch.cuda.device(E.device):
                _lib.check(h.ge2e_b200_embed_tail_bwd_gemms(dY.data_ptr(), w.data_ptr(), xv.data_ptr(), xv.stride(0), U, H, D,
                                                            dx_last.data_ptr() if want_x else None, dxs,
                                                            dW.data_ptr() if want_w else None, ops._stream()),
                           "ge2e_b200_embed_tail_bwd_gemms")
        else:                                                       # H % 32 != 0: plain library GEMMs
            if want_x:
                dx_last.copy_(dY @ w)
            if want_w:
                torch.matmul(dY.t(), xv, out=dW)
        return dX, dW, dbias


def project_normalize(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """normalise_rows(last_frame(x) @ weight.T + bias); ``x`` is [U, H] or the LSTM output [U, frames, H]."""
    return _ProjectNormalize.apply(x, weight, bias)


class ProjectionL2Norm(nn.Module):
    def __init__(self, hidden_size: int, embedding_size: int):
        super().__init__()
        self.projection = nn.Linear(hidden_size, embedding_size)    # s2:25

    def forward(self, x):
        return project_normalize(x, self.projection.weight, self.projection.bias)
