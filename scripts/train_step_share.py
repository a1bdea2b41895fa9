"""BASELINE config 5: one synthetic training step of the reference's embedder shape (3-layer LSTM,
40 mel x 160 frames, 768 hidden, 256-d projection, L2-normalised output; N=64 speakers x M=10
utterances) with the step sequence of s4_train_embed_model.py:188-205, and the share of that step
spent in the GE2E loss: share = t(loss fwd+bwd alone, same shapes, eager module API) / t(step),
for the CUDA loss of this repo and for an eager PyTorch GE2E written the way the reference does it
(expanded cosine similarity), both on the GPU.  Inputs are synthetic, weights random.
    python scripts/train_step_share.py [N M]"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speaker_embedding_ge2e_loss_b200 import GE2ELoss  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T, MEL, HID, EMB, LAYERS = 160, 40, 768, 256, 3
dev = torch.device("cuda:0")
torch.manual_seed(0)


class Embedder(nn.Module):
    """Shape of the reference's ModelGE2ELossSpeachEmbed (s2:13-35): LSTM -> last frame -> Linear -> L2 norm."""

    def __init__(self):
        super().__init__()
        self.lstm = nn.LSTM(MEL, HID, num_layers=LAYERS, batch_first=True)
        self.proj = nn.Linear(HID, EMB)

    def forward(self, x):
        y, _ = self.lstm(x.float())
        e = self.proj(y[:, -1].float())
        return e / e.norm(dim=1, keepdim=True)


class EagerGE2E(nn.Module):
    """Plain PyTorch softmax GE2E in the reference's formulation (expanded rows, F.cosine_similarity)."""

    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.tensor(10.0, device=dev))
        self.b = nn.Parameter(torch.tensor(-5.0, device=dev))

    def forward(self, E):
        n, m, d = E.shape
        c = E.mean(1)
        u = (E.sum(1, keepdim=True) - E) / (m - 1)
        same = F.cosine_similarity(E.reshape(-1, d), u.reshape(-1, d))
        cr = c.repeat(n * m, 1)
        er = E.reshape(-1, d).unsqueeze(1).repeat(1, n, 1).reshape(-1, d)
        cos = F.cosine_similarity(er, cr).view(n, m, n)
        idx = torch.arange(n, device=E.device)
        cos[idx, :, idx] = same.view(n, m)
        S = self.w * (cos + 1e-6) + self.b
        return (torch.log(torch.exp(S).sum(2) + 1e-6) - S[idx, :, idx]).sum()


class NullLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.tensor(10.0, device=dev))
        self.b = nn.Parameter(torch.tensor(-5.0, device=dev))

    def forward(self, E):
        return E.sum() * 1e-3 + 0.0 * (self.w + self.b)


def timed(crit, steps=12, warmup=4):
    model = Embedder().to(dev)
    opt = torch.optim.SGD([{"params": model.parameters()}, {"params": crit.parameters()}], lr=0.01)
    x = torch.rand(N * M, T, MEL, device=dev)
    perm = torch.randperm(N * M, device=dev)
    unperm = torch.argsort(perm)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for k in range(warmup + steps):
        if k == warmup:
            torch.cuda.synchronize()
            ev[0].record()
        emb = model(x[perm])[unperm].reshape(N, M, EMB)              # s4:174-192
        loss = crit(emb)                                             # s4:196
        opt.zero_grad()
        loss.backward()                                              # s4:200
        torch.nn.utils.clip_grad_norm_(model.parameters(), 3.0)      # s4:201
        torch.nn.utils.clip_grad_norm_(crit.parameters(), 1.0)       # s4:202
        opt.step()
        loss.to("cpu").detach().numpy()                              # s4:205 (host sync every step)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / steps


def loss_alone(crit, steps=50, warmup=10):
    """fwd + bwd of the loss module alone on embeddings of the same shape (device time, eager)."""
    E = F.normalize(torch.randn(N, M, EMB, device=dev), dim=2).requires_grad_(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for k in range(warmup + steps):
        if k == warmup:
            torch.cuda.synchronize()
            ev[0].record()
        E.grad = None
        crit(E).backward()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / steps


if __name__ == "__main__":
    print(f"N={N} M={M}  embedder: LSTM {MEL}->{HID}x{LAYERS} -> Linear {EMB}, {T} frames")
    for name, make in (("this repo, fp32", lambda: GE2ELoss(None, device=dev)),
                       ("this repo, tf32", lambda: GE2ELoss(None, device=dev, precision="tf32")),
                       ("eager PyTorch  ", EagerGE2E)):
        t_step = timed(make())
        t_loss = loss_alone(make())
        print(f"  {name}: step {t_step:8.3f} ms   loss fwd+bwd alone {t_loss * 1e3:8.1f} us   share {t_loss / t_step * 100:5.2f} %")
    print(f"  null loss      : step {timed(NullLoss()):8.3f} ms")
