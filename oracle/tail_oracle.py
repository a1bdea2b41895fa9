"""TEST INFRASTRUCTURE ONLY -- fp64 numpy restatement of the reference model's tail.

Follows /root/reference/embedding_model_GE2E/s2_model_GE2E_loss_speach_embed.py:
  * :30  ``x = x[:, x.size(1) - 1]``            last frame of the LSTM output [U, frames, H]
  * :31  ``x = self.projection(x.float())``     nn.Linear(H, D): x W^T + bias  (:25)
  * :34  ``x = x / torch.norm(x, dim=1).unsqueeze(1)``   L2 normalise, no epsilon
and the analytic backward of those three lines.  Pinned against the reference's own module (real LSTM +
projection, torch autograd) by tests/golden/make_tail_golden.py -> tests/golden/tail_reference_vectors.npz.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np


def embed_tail(x_last, W, bias):
    """(E[U, D] unit rows, inv_norm[U]) from the last-frame activations x_last[U, H]."""
    x = np.asarray(x_last, dtype=np.float64)
    y = x @ np.asarray(W, dtype=np.float64).T
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)
    n = np.sqrt((y * y).sum(axis=1))
    return y / n[:, None], 1.0 / n


def embed_tail_backward(x_last, W, bias, dE):
    """Gradients of sum(E * dE) wrt x_last, W, bias."""
    x = np.asarray(x_last, dtype=np.float64)
    Wd = np.asarray(W, dtype=np.float64)
    g = np.asarray(dE, dtype=np.float64)
    e, inv = embed_tail(x, Wd, bias)
    dY = (g - e * (e * g).sum(axis=1, keepdims=True)) * inv[:, None]
    return dict(dX=dY @ Wd, dW=dY.T @ x, dbias=dY.sum(axis=0), dY=dY)
