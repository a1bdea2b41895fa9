"""Generate tests/golden/batch_reference_vectors.npz with the REAL reference dataset class
(EmbeddingModelTTDataset.__getitem__, s1:52-79) on seeded ``sv_*.npy`` files written to a temporary
directory, plus the trainer's reshape / perm / float lines (s4:170-186, s2:28) applied with torch.
Run only where /root/reference is mounted:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_batch_golden.py
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from embedding_model_GE2E.s1_dataset_loader import EmbeddingModelTTDataset  # noqa: E402
from utils.dict_to_dot import GetDictWithDotNotation  # noqa: E402


def main():
    out = {}
    for name, S, frames, mels, N, M, L, seed in [("s6_n4m5", 6, 180, 40, 4, 5, 160, 1), ("s5_n5m3_small", 5, 30, 8, 5, 3, 20, 2),
                                                   ("s9_n8m4_odd", 9, 45, 6, 8, 4, 31, 3)]:
        rng = np.random.default_rng(seed)
        with tempfile.TemporaryDirectory() as d:
            files = []
            for s in range(S):
                utts = int(rng.integers(3, 9))
                arr = rng.standard_normal((utts, frames, mels))            # float64, as s0 saves it
                np.save(os.path.join(d, f"sv_spk{s:02d}.npy"), arr)
                files.append(arr)
            hp = GetDictWithDotNotation({"m_ge2e": {"training_M": M, "test_M": M,
                                                    "tt_data": {"min_train_utter_len": L, "min_test_utter_len": L}}})
            ds = EmbeddingModelTTDataset(data_path=d, hp=hp, training=False)     # no list shuffle: os.walk order
            order = [int(f[6:8]) for f in ds.lst_spkr_np_files]                 # which array each index loads
            np.random.seed(seed)
            idx = list(range(N))
            items = [ds[i] for i in idx]                                         # the DataLoader's calls, in order
        batch = torch.tensor(np.stack(items))                                    # default collate: float64 [N, M, L, mels]
        flat = torch.reshape(batch, (N * M, batch.shape[2], batch.shape[3]))     # s4:176-177
        random.seed(seed)
        perm = random.sample(range(0, N * M), N * M)                             # s4:179
        model_input = flat[perm].float()                                         # s4:186, s2:28
        for s in range(S):
            out[f"{name}_file{s}"] = files[s]
        out[f"{name}_order"] = np.asarray(order[:N], dtype=np.int64)
        out[f"{name}_cfg"] = np.asarray([S, frames, mels, N, M, L, seed], dtype=np.int64)
        out[f"{name}_perm"] = np.asarray(perm, dtype=np.int64)
        out[f"{name}_batch"] = model_input.numpy()
        print(name, tuple(model_input.shape), order[:N])
    np.savez_compressed(os.path.join(HERE, "batch_reference_vectors.npz"), **out)


if __name__ == "__main__":
    main()
