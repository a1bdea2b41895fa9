"""torch.library custom ops over the C ABI (CUDA only; raises on CPU tensors).

``ge2e_b200::fwd`` / ``ge2e_b200::bwd`` are the single-device ops; the staged ops
(``prep``, ``fwd_rows``, ``bwd_rows``, ``bwd_finalize``) are what the speaker-sharded
autograd function in ``sharded.py`` strings together around its collectives.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts: Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("speaker_embedding_ge2e_loss_b200 runs on CUDA (sm_100a) tensors only; "
                               "there is no CPU fallback")


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t.contiguous()


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _workspace(n_local: int, n_total: int, M: int, D: int, variant: int, precision: int, device):
    nbytes = lib().ge2e_b200_workspace_bytes(n_local, n_total, M, D, variant, precision)
    if nbytes == 0:
        return None, 0
    # contract of the C ABI: zero on entry; the kernels leave it zeroed, so a caller that keeps the
    # buffer (GE2EPlan) zeroes it once
    return torch.zeros(nbytes, dtype=torch.uint8, device=device), nbytes


# --------------------------------------------------------------------------- single device
@torch.library.custom_op("ge2e_b200::fwd", mutates_args=())
def ge2e_fwd(E: Tensor, w: Tensor, b: Tensor, eps: float, variant: int,
             precision: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """GE2ELoss.forward (reference s3:19-30).  Returns (loss, e_hat, c_hat, cos_diag, row_stat,
    row_kstar, row_aux); everything after loss is saved for the backward."""
    _need_cuda(E, w, b)
    E, w, b = _f32c(E), _f32c(w), _f32c(b)
    N, M, D = E.shape
    U = N * M
    dev = E.device
    with torch.cuda.device(dev):
        accum = torch.empty(4, dtype=torch.float32, device=dev)
        e_hat = torch.empty((U, D), dtype=torch.float32, device=dev)
        c_hat = torch.empty((N, D), dtype=torch.float32, device=dev)
        cos_diag = torch.empty(U, dtype=torch.float32, device=dev)
        row_stat = torch.empty(U, dtype=torch.float32, device=dev)
        row_kstar = torch.empty(U if variant == _lib.CONTRAST else 1, dtype=torch.int32, device=dev)
        row_aux = torch.empty(U, dtype=torch.float32, device=dev)
        ws, ws_bytes = _workspace(N, N, M, D, variant, precision, dev)
        rc = lib().ge2e_b200_forward(E.data_ptr(), N, M, D, w.data_ptr(), b.data_ptr(), eps, variant,
                                     precision, e_hat.data_ptr(), c_hat.data_ptr(), cos_diag.data_ptr(),
                                     row_stat.data_ptr(), row_kstar.data_ptr(), row_aux.data_ptr(),
                                     accum.data_ptr(), _ptr(ws), ws_bytes, _stream())
    check(rc, "ge2e_b200_forward")
    return accum[0], e_hat, c_hat, cos_diag, row_stat, row_kstar, row_aux


@ge2e_fwd.register_fake
def _(E, w, b, eps, variant, precision):
    N, M, D = E.shape
    U = N * M
    return (E.new_empty(()), E.new_empty((U, D)), E.new_empty((N, D)), E.new_empty(U), E.new_empty(U),
            E.new_empty(U if variant == _lib.CONTRAST else 1, dtype=torch.int32), E.new_empty(U))


@torch.library.custom_op("ge2e_b200::bwd", mutates_args=())
def ge2e_bwd(grad_out: Tensor, E: Tensor, w: Tensor, b: Tensor, e_hat: Tensor, c_hat: Tensor,
             cos_diag: Tensor, row_stat: Tensor, row_kstar: Tensor, row_aux: Tensor, eps: float, variant: int,
             precision: int) -> Tuple[Tensor, Tensor]:
    """Backward of ge2e_b200::fwd.  Returns (dE[N,M,D], dwdb[2])."""
    _need_cuda(grad_out, E, w, b)
    E, w, b = _f32c(E), _f32c(w), _f32c(b)
    g = _f32c(grad_out)
    N, M, D = E.shape
    U = N * M
    dev = E.device
    with torch.cuda.device(dev):
        dE = torch.empty_like(E)
        dE_hat = torch.empty((U, D), dtype=torch.float32, device=dev)
        # dC_hat followed by {dw, db}: the library zeroes both with one memset; the layout is
        # [dC_hat (N*D) | dw | db] and `accum` = start of (dw - 1) so that accum[1]=dw, accum[2]=db
        scratch = torch.empty(N * D + 2, dtype=torch.float32, device=dev)
        ws, ws_bytes = _workspace(N, N, M, D, variant, precision, dev)
        accum_ptr = scratch.data_ptr() + (N * D - 1) * 4
        rc = lib().ge2e_b200_backward(E.data_ptr(), e_hat.data_ptr(), c_hat.data_ptr(), cos_diag.data_ptr(),
                                      row_stat.data_ptr(), row_kstar.data_ptr(), row_aux.data_ptr(), N, M, D,
                                      w.data_ptr(), b.data_ptr(), eps, variant, precision, g.data_ptr(),
                                      dE_hat.data_ptr(), scratch.data_ptr(), accum_ptr, dE.data_ptr(),
                                      _ptr(ws), ws_bytes, _stream())
    check(rc, "ge2e_b200_backward")
    return dE, scratch[N * D:]


@ge2e_bwd.register_fake
def _(grad_out, E, w, b, e_hat, c_hat, cos_diag, row_stat, row_kstar, row_aux, eps, variant, precision):
    return torch.empty_like(E), E.new_empty(2)


def _fwd_setup(ctx, inputs, output):
    E, w, b, eps, variant, precision = inputs
    _, e_hat, c_hat, cos_diag, row_stat, row_kstar, row_aux = output
    ctx.save_for_backward(E, w, b, e_hat, c_hat, cos_diag, row_stat, row_kstar, row_aux)
    ctx.cfg = (eps, variant, precision)


def _fwd_backward(ctx, g_loss, *_unused):
    E, w, b, e_hat, c_hat, cos_diag, row_stat, row_kstar, row_aux = ctx.saved_tensors
    eps, variant, precision = ctx.cfg
    dE, dwdb = ge2e_bwd(g_loss, E, w, b, e_hat, c_hat, cos_diag, row_stat, row_kstar, row_aux, eps, variant,
                        precision)
    return dE, dwdb[0], dwdb[1], None, None, None


ge2e_fwd.register_autograd(_fwd_backward, setup_context=_fwd_setup)


def ge2e_loss(E: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6, variant: str = "softmax",
              precision: str = "fp32") -> Tensor:
    """Functional form: differentiable in E, w, b."""
    if E.dim() != 3:
        raise ValueError(f"embeddings must be [N, M, D], got {tuple(E.shape)}")
    if E.shape[1] < 2:
        raise ValueError("GE2E needs M >= 2 utterances per speaker (the reference divides by M - 1)")
    return ge2e_fwd(E, w, b, float(eps), _lib.VARIANTS[variant], _lib.PRECISIONS[precision])[0]


# --------------------------------------------------------------------------- staged (plain functions)
def prep(E: Tensor, c_hat_local_out: Tensor, precision: int):
    """Stage 1 on the local speakers; writes c_hat into ``c_hat_local_out`` (a slice of the
    all-gather buffer).  Returns (e_hat, cos_diag, accum)."""
    _need_cuda(E, c_hat_local_out)
    n, M, D = E.shape
    dev = E.device
    e_hat = torch.empty((n * M, D), dtype=torch.float32, device=dev)
    cos_diag = torch.empty(n * M, dtype=torch.float32, device=dev)
    accum = torch.empty(4, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib().ge2e_b200_prep(E.data_ptr(), n, M, D, precision, e_hat.data_ptr(),
                                  c_hat_local_out.data_ptr(), cos_diag.data_ptr(), accum.data_ptr(), _stream())
    check(rc, "ge2e_b200_prep")
    return e_hat, cos_diag, accum


def fwd_rows(e_hat, c_hat_all, cos_diag, n_local, n_total, spk_offset, M, D, w, b, eps, variant, precision,
             accum, per_row: bool = False, sim: bool = False):
    dev = e_hat.device
    U = n_local * M
    row_stat = torch.empty(U, dtype=torch.float32, device=dev)
    row_kstar = torch.empty(U if variant == _lib.CONTRAST else 1, dtype=torch.int32, device=dev)
    row_aux = torch.empty(U, dtype=torch.float32, device=dev)
    per = torch.empty(U, dtype=torch.float32, device=dev) if per_row else None
    sim_out = torch.empty((U, n_total), dtype=torch.float32, device=dev) if sim else None
    with torch.cuda.device(dev):
        ws, ws_bytes = _workspace(n_local, n_total, M, D, variant, precision, dev)
        rc = lib().ge2e_b200_fwd_rows(e_hat.data_ptr(), c_hat_all.data_ptr(), cos_diag.data_ptr(), n_local,
                                      n_total, spk_offset, M, D, w.data_ptr(), b.data_ptr(), eps, variant,
                                      precision, row_stat.data_ptr(), row_kstar.data_ptr(), row_aux.data_ptr(),
                                      accum.data_ptr(), _ptr(per), _ptr(sim_out), _ptr(ws), ws_bytes, _stream())
    check(rc, "ge2e_b200_fwd_rows")
    return row_stat, row_kstar, row_aux, per, sim_out


def bwd_rows(e_hat, c_hat_all, cos_diag, row_stat, row_kstar, row_aux, n_local, n_total, spk_offset, M, D, w, b,
             eps, variant, precision, grad_out):
    """Returns (dE_hat[U_local, D], dC_hat_partial[n_total, D], dwdb[2])."""
    dev = e_hat.device
    dE_hat = torch.empty((n_local * M, D), dtype=torch.float32, device=dev)
    scratch = torch.empty(n_total * D + 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws, ws_bytes = _workspace(n_local, n_total, M, D, variant, precision, dev)
        rc = lib().ge2e_b200_bwd_rows(e_hat.data_ptr(), c_hat_all.data_ptr(), cos_diag.data_ptr(),
                                      row_stat.data_ptr(), row_kstar.data_ptr(), row_aux.data_ptr(), n_local,
                                      n_total, spk_offset, M, D, w.data_ptr(), b.data_ptr(), eps, variant, precision,
                                      grad_out.data_ptr(), dE_hat.data_ptr(), scratch.data_ptr(),
                                      scratch.data_ptr() + n_total * D * 4, _ptr(ws), ws_bytes, _stream())
    check(rc, "ge2e_b200_bwd_rows")
    return dE_hat, scratch[:n_total * D].view(n_total, D), scratch[n_total * D:]


def bwd_finalize(E, dE_hat, dC_hat_local, cos_diag, row_stat, row_aux, w, b, eps, variant, grad_out):
    n, M, D = E.shape
    dE = torch.empty_like(E)
    with torch.cuda.device(E.device):
        rc = lib().ge2e_b200_bwd_finalize(E.data_ptr(), dE_hat.data_ptr(), dC_hat_local.data_ptr(),
                                          cos_diag.data_ptr(), row_stat.data_ptr(), row_aux.data_ptr(), n, M, D,
                                          w.data_ptr(),
                                          b.data_ptr(), eps, variant, grad_out.data_ptr(), dE.data_ptr(),
                                          _stream())
    check(rc, "ge2e_b200_bwd_finalize")
    return dE
