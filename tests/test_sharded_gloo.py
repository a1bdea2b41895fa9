"""world_size-2/4 gloo tests of the speaker-sharded plumbing (all-gather of C_hat, reduce-scatter of
dC_hat, all-reduce of loss / dw / db) on CPU.  The four compute stages are replaced by a torch
fp64 test double with the C-ABI stages' exact contract (include/ge2e_b200.h); the result must
equal the single-process oracle on the concatenated batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DELTA = 1e-8


class TorchStages:
    """CPU stand-in for ops.prep / fwd_rows / bwd_rows / bwd_finalize (same signatures).  ``fused``: emulate
    the tensor-core softmax contract (the forward leaves un-normalised dE_hat rows + row_scale)."""
    fused = False

    @staticmethod
    def _unit(x):
        n = x.norm(dim=-1, keepdim=True)
        return x / n.clamp_min(DELTA), n

    @staticmethod
    def prep(E, c_hat_local_out, precision):
        n, M, D = E.shape
        s = E.sum(dim=1, keepdim=True)
        u = (s - E) / (M - 1)
        eh, _ = TorchStages._unit(E.reshape(n * M, D))
        uh, _ = TorchStages._unit(u.reshape(n * M, D))
        ch, _ = TorchStages._unit(s[:, 0] / M)
        c_hat_local_out.copy_(ch)
        return eh, (eh * uh).sum(-1), torch.zeros(4, dtype=E.dtype)

    @staticmethod
    def _S(e_hat, c_hat_all, cos_diag, n_local, spk_offset, M, w, b, eps):
        U = n_local * M
        cos = e_hat @ c_hat_all.T
        rows = torch.arange(U)
        spk = spk_offset + rows // M
        cos[rows, spk] = cos_diag
        return w * (cos + eps) + b, cos + eps, rows, spk

    @staticmethod
    def fwd_rows(e_hat, c_hat_all, cos_diag, n_local, n_total, spk_offset, M, D, w, b, eps, variant, precision,
                 accum, per_row=False, sim=False, want_grad=False):
        S, _, rows, spk = TorchStages._S(e_hat, c_hat_all, cos_diag, n_local, spk_offset, M, w, b, eps)
        aux = torch.zeros_like(cos_diag)
        dE_hat = row_scale = None
        if variant == 0:
            Z = torch.exp(S).sum(1) + eps
            stat = torch.log(Z)
            per = stat - S[rows, spk]
            aux = (Z - torch.exp(S[rows, spk])) / Z          # q = 1 - p_jj
            kstar = torch.zeros(1, dtype=torch.int32)
            if want_grad and TorchStages.fused:
                # contract of the tensor-core rows pass (include/ge2e_b200.h, ge2e_b200_fwd_rows): P against
                # the fixed shift |w| + (w eps + b), own-speaker column excluded; un-normalised rows + scale
                m = w.abs() + (w * eps + b)
                P = torch.exp(S - m)
                P[rows, spk] = 0
                dE_hat = P @ c_hat_all
                row_scale = w * torch.exp(m - stat)
        else:
            Sm = S.clone()
            Sm[rows, spk] = -float("inf")
            stat, kstar = Sm.max(dim=1)
            per = 1 - torch.sigmoid(S[rows, spk]) + torch.sigmoid(stat)
        accum[0] += per.sum()
        return stat, kstar, aux, per, None, dE_hat, row_scale

    @staticmethod
    def bwd_rows(e_hat, c_hat_all, cos_diag, row_stat, row_kstar, row_aux, n_local, n_total, spk_offset, M, D, w,
                 b, eps, variant, precision, grad_out, dE_hat=None, row_scale=None):
        S, cos, rows, spk = TorchStages._S(e_hat, c_hat_all, cos_diag, n_local, spk_offset, M, w, b, eps)
        if variant == 0:
            G = torch.exp(S - row_stat[:, None])
            G[rows, spk] = -row_aux
        else:
            G = torch.zeros_like(S)
            sp = torch.sigmoid(S[rows, spk])
            G[rows, spk] = -sp * (1 - sp)
            sn = torch.sigmoid(row_stat)
            G[rows, row_kstar.long()] += sn * (1 - sn)
        G = G * grad_out
        dwdb = torch.stack([(G * cos).sum(), G.sum()])
        Goff = (w * G).clone()
        Goff[rows, spk] = 0
        if row_scale is not None:                      # the forward prepared dE_hat: centroid pass only
            return dE_hat, Goff.T @ e_hat, dwdb
        return Goff @ c_hat_all, Goff.T @ e_hat, dwdb

    @staticmethod
    def bwd_finalize(E, dE_hat, dC_hat_local, cos_diag, row_stat, row_aux, w, b, eps, variant, grad_out,
                     row_scale=None):
        n, M, D = E.shape
        if row_scale is not None:
            dE_hat = dE_hat * (grad_out * row_scale)[:, None]
        Ef = E.reshape(n * M, D)
        s = E.sum(dim=1, keepdim=True)
        u = ((s - E) / (M - 1)).reshape(n * M, D)
        eh, ne = TorchStages._unit(Ef)
        uh, nu = TorchStages._unit(u)
        ch, nc = TorchStages._unit(s[:, 0] / M)
        Sd = w * (cos_diag + eps) + b
        if variant == 0:
            Gd = -grad_out * row_aux
        else:
            sp = torch.sigmoid(Sd)
            Gd = -grad_out * sp * (1 - sp)
        dd = (w * Gd)[:, None]

        def unit_bwd(dxh, xh, nrm):
            return (dxh - xh * (xh * dxh).sum(-1, keepdim=True)) / nrm.clamp_min(DELTA)

        de = unit_bwd(dE_hat + dd * uh, eh, ne).reshape(n, M, D)
        du = unit_bwd(dd * eh, uh, nu).reshape(n, M, D)
        dc = unit_bwd(dC_hat_local, ch, nc)
        return de + dc[:, None, :] / M + (du.sum(1, keepdim=True) - du) / (M - 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, M, D, variant, q, fused=False):
    TorchStages.fused = fused
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ge2e_oracle as orc
        from speaker_embedding_ge2e_loss_b200.sharded import shard_bounds, sharded_ge2e_loss
        torch.set_num_threads(1)
        E_all = torch.tensor(orc.make_embeddings(N, M, D, seed=7, kind="clustered"), dtype=torch.float64)
        off, n_local = shard_bounds(N, world, rank)
        E = E_all[off:off + n_local].clone().requires_grad_(True)
        w = torch.tensor(10.0, dtype=torch.float64, requires_grad=True)
        b = torch.tensor(-5.0, dtype=torch.float64, requires_grad=True)
        loss = sharded_ge2e_loss(E, w, b, 1e-6, variant, "fp32", None, stages=TorchStages)
        (loss * 0.5).backward()
        q.put((rank, loss.item(), E.grad.numpy(), w.grad.item(), b.grad.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,variant,fused", [(2, "softmax", False), (2, "softmax", True), (2, "contrast", False),
                                                 (4, "softmax", True)])
def test_sharded_equals_single_process_oracle(world, variant, fused):
    from oracle import ge2e_oracle as orc
    N, M, D = 8, 3, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, M, D, variant, q, fused)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = orc.forward_backward(orc.make_embeddings(N, M, D, seed=7, kind="clustered"), 10.0, -5.0, 1e-6, variant,
                               g=0.5)
    dE = np.concatenate([r[2] for r in res], axis=0)
    for r in res:                                   # every rank sees the global loss, dw, db
        assert abs(r[1] - ref["loss"]) < 1e-9 * max(1, abs(ref["loss"]))
        assert abs(r[3] - ref["dw"]) < 1e-9 * max(1, abs(ref["dw"]))
        assert abs(r[4] - ref["db"]) < 1e-9
    assert np.linalg.norm(dE - ref["dE"]) < 1e-9 * np.linalg.norm(ref["dE"])


def test_shard_bounds():
    from speaker_embedding_ge2e_loss_b200.sharded import shard_bounds
    assert shard_bounds(8192, 8, 3) == (3072, 1024)
    assert shard_bounds(8, 1, 0) == (0, 8)
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
