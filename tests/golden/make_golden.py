"""Generate tests/golden/*.npz by running the REAL reference GE2ELoss (read-only import
from /root/reference) in this container.  /root/reference does not exist on the GPU box,
so the vectors are committed; re-run this script only where the reference is mounted:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

For every case it stores the input, and loss / per-embedding loss / dE / dw / db from
 (a) the reference in fp32 (its native precision, what a user of the reference sees) and
 (b) the reference in fp64 (same code; dtype follows the input) = "truth".
The contrast variant does not exist in the reference: its fixtures take S from the
reference's own ``get_cos_sim`` (s3:41-80) and apply eq. (7) of arXiv:1710.10467 on top
with torch autograd in fp64.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from embedding_model_GE2E.s3_loss_function_GE2E import GE2ELoss as RefLoss  # noqa: E402
from utils.dict_to_dot import GetDictWithDotNotation  # noqa: E402

from oracle.ge2e_oracle import make_embeddings  # noqa: E402

EPS = 1e-6
HP = GetDictWithDotNotation({"general": {"device": torch.device("cpu"), "small_err": EPS}})


def run_ref(E_np, w, b, dtype):
    crit = RefLoss(HP)
    with torch.no_grad():
        crit.w.data = torch.tensor(w, dtype=dtype)
        crit.b.data = torch.tensor(b, dtype=dtype)
    E = torch.tensor(E_np, dtype=dtype, requires_grad=True)
    loss = crit(E)
    loss.backward()
    with torch.no_grad():
        S = crit.w * RefLoss.get_cos_sim(E, RefLoss.get_centroids(E), HP) + crit.b
        _, per = RefLoss.calc_loss(S, HP)
    return dict(loss=loss.item(), per=per.numpy(), dE=E.grad.numpy(),
                dw=crit.w.grad.item(), db=crit.b.grad.item())


def run_contrast(E_np, w, b):
    E = torch.tensor(E_np, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    S = wt * RefLoss.get_cos_sim(E, RefLoss.get_centroids(E), HP) + bt
    N = S.shape[0]
    j = list(range(N))
    pos = torch.sigmoid(S[j, :, j])
    per = 1.0 - pos
    if N > 1:
        mask = torch.zeros_like(S, dtype=torch.bool)
        mask[j, :, j] = True
        neg = torch.sigmoid(S.masked_fill(mask, -float("inf"))).max(dim=2).values
        per = per + neg
    loss = per.sum()
    loss.backward()
    return dict(loss=loss.item(), per=per.detach().numpy(), dE=E.grad.numpy(),
                dw=wt.grad.item(), db=bt.grad.item())


CASES = [
    # name, N, M, D, kind, seed, w, b
    ("cfg1_4x8x256_random", 4, 8, 256, "random", 0, 10.0, -5.0),
    ("repo_2x16x256_random", 2, 16, 256, "random", 1, 10.0, -5.0),
    ("cfg2_64x10x256_random", 64, 10, 256, "random", 2, 10.0, -5.0),
    ("cfg2_64x10x256_clustered", 64, 10, 256, "clustered", 3, 10.0, -5.0),
    ("raw_5x3x32", 5, 3, 32, "raw", 4, 7.5, -2.0),
    ("negw_3x2x8", 3, 2, 8, "random", 5, -3.0, 1.0),
    ("onespk_1x4x16", 1, 4, 16, "random", 6, 10.0, -5.0),
    ("ragged_d_7x5x100", 7, 5, 100, "clustered", 7, 10.0, -5.0),
    ("n130_130x2x64", 130, 2, 64, "random", 8, 10.0, -5.0),
]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)  # deterministic reduction order for the fp32 fixtures
    out = {}
    # KAT input of s3:144-145 (the only fixed vector in the reference; outputs derived here)
    kat = np.array([[0, 1, 0], [0, 0, 1], [0, 1, 0], [0, 1, 0], [1, 0, 0], [1, 0, 0]],
                   dtype=np.float32).reshape(3, 2, 3)
    for tag, (w, b) in {"kat_w1_b0": (1.0, 0.0), "kat_w10_bm5": (10.0, -5.0)}.items():
        r32 = run_ref(kat, w, b, torch.float32)
        r64 = run_ref(kat, w, b, torch.float64)
        rc = run_contrast(kat, w, b)
        out[tag] = dict(E=kat, w=w, b=b, r32=r32, r64=r64, rc=rc)
    for name, N, M, D, kind, seed, w, b in CASES:
        E = make_embeddings(N, M, D, seed=seed, kind=kind)
        out[name] = dict(E=E, w=w, b=b, r32=run_ref(E, w, b, torch.float32),
                         r64=run_ref(E, w, b, torch.float64), rc=run_contrast(E, w, b))
        print(name, out[name]["r32"]["loss"], out[name]["r64"]["loss"], out[name]["rc"]["loss"])

    flat = {}
    for name, c in out.items():
        flat[f"{name}/E"] = c["E"].astype(np.float32)
        flat[f"{name}/wb"] = np.array([c["w"], c["b"]], dtype=np.float64)
        for tag in ("r32", "r64", "rc"):
            r = c[tag]
            flat[f"{name}/{tag}/scalars"] = np.array([r["loss"], r["dw"], r["db"]], dtype=np.float64)
            flat[f"{name}/{tag}/per"] = r["per"].astype(np.float64)
            # fp64 results are stored as fp32 to keep the fixture small (6e-8 relative)
            flat[f"{name}/{tag}/dE"] = r["dE"].astype(np.float32)
    path = os.path.join(HERE, "ge2e_reference_vectors.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB; torch", torch.__version__)


if __name__ == "__main__":
    main()
