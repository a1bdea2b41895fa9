"""Generate tests/golden/eer_reference_vectors.npz by executing the REFERENCE'S OWN SOURCE TEXT of the
EER sweep (s5_eval_model.py, from ``diff = 1`` to the line before the final print; the function around
it needs a trained model and a data loader, so the block is cut out of the file by its markers and
exec'd on seeded similarity matrices; per-threshold FAR / FRR come from the same block with its
``thres_lst`` line narrowed to one threshold).  Run only where /root/reference is mounted:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_eer_golden.py
"""
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
SRC = "/root/reference/embedding_model_GE2E/s5_eval_model.py"

sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
import torch  # noqa: E402
from embedding_model_GE2E.s3_loss_function_GE2E import GE2ELoss as RefLoss  # noqa: E402
from utils.dict_to_dot import GetDictWithDotNotation  # noqa: E402

from oracle.ge2e_oracle import make_embeddings  # noqa: E402

HP = GetDictWithDotNotation({"general": {"device": torch.device("cpu"), "small_err": 1e-6}})


def reference_block():
    lines = open(SRC).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip() == "diff = 1")
    end = next(i for i, l in enumerate(lines) if l.strip().startswith('print("\\nEER'))
    return textwrap.dedent("\n".join(lines[start:end]))


def overlapping_speakers(N, M, D, seed, common, spread):
    """Unit embeddings whose own-speaker and cross-speaker cosines both straddle the 0.5..0.99 sweep:
    a component shared by all speakers (cross-speaker similarity), one per speaker, and noise."""
    rng = np.random.default_rng(seed)
    shared = rng.standard_normal(D)
    shared /= np.linalg.norm(shared)
    centres = rng.standard_normal((N, D))
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    noise = rng.standard_normal((N, M, D)) / np.sqrt(D)
    E = common * shared[None, None, :] + centres[:, None, :] + spread * noise
    return (E / np.linalg.norm(E, axis=2, keepdims=True)).astype(np.float32)


def cos_matrix(E_np):
    """float32 similarity matrix exactly as s5:42-46 builds it: the reference's own get_centroids /
    get_cos_sim on float32 embeddings, w = 1, b = 0."""
    E = torch.tensor(E_np, dtype=torch.float32)
    with torch.no_grad():
        cos = RefLoss.get_cos_sim(E, RefLoss.get_centroids(E), HP)
        S = torch.tensor(1.0) * cos + torch.tensor(0.0)
    return S.detach().cpu().numpy()


def main():
    block = reference_block()
    code = compile(block, SRC + ":EER-block", "exec")
    out = {}
    cases = [("n4m5", 4, 5, 64, 1, 1.0, 1.0), ("n4m6", 4, 6, 256, 2, 1.5, 0.8), ("n8m5", 8, 5, 32, 3, 2.0, 1.2),
             ("n16m10", 16, 10, 256, 4, 1.2, 1.5), ("n2m3", 2, 3, 16, 5, 0.8, 0.6), ("n32m4", 32, 4, 24, 6, 1.8, 0.9),
             ("n6m7_sep", 6, 7, 128, 7, 0.0, 0.3), ("n5m4_noisy", 5, 4, 40, 8, 3.0, 2.5)]
    for name, N, M, D, seed, common, spread in cases:
        S = cos_matrix(overlapping_speakers(N, M, D, seed, common, spread))
        hp = types.SimpleNamespace(m_ge2e=types.SimpleNamespace(test_N=N, test_M=M))
        ns = {"S": S, "hp": hp, "np": np}
        exec(code, ns)
        out[name + "_S"] = S
        # per-threshold FAR / FRR: the same block with only its threshold list narrowed to one value
        thr_line = next(l for l in block.split("\n") if l.strip().startswith("thres_lst = "))
        far, frr = [], []
        for th in [0.01 * i + 0.5 for i in range(50)]:
            one = block.replace(thr_line, thr_line[:len(thr_line) - len(thr_line.lstrip())] + f"thres_lst = [{th!r}]")
            ns1 = {"S": S, "hp": hp, "np": np}
            exec(compile(one, SRC + ":EER-block-1", "exec"), ns1)
            far.append(ns1["FAR"])
            frr.append(ns1["FRR"])
        out[name + "_far"] = np.array(far, dtype=np.float64)
        out[name + "_frr"] = np.array(frr, dtype=np.float64)
        out[name + "_res"] = np.array([ns["EER"], ns["EER_thres"], ns["EER_FAR"], ns["EER_FRR"]], dtype=np.float64)
        print(name, S.shape, out[name + "_res"])
    np.savez_compressed(os.path.join(HERE, "eer_reference_vectors.npz"), **out)


if __name__ == "__main__":
    main()
