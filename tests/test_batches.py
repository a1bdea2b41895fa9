"""Batch assembly (s1_dataset_loader.py:52-79 + s4_train_embed_model.py:170-186).

CPU: the oracle restatement against batches produced by the reference's own dataset class
(tests/golden/make_batch_golden.py).  GPU: SpectrogramBank + the gather kernel, bit-exact against the
goldens (same seed -> same crops -> same float32 batch) and against the oracle on seeded banks.
"""
import os
import random

import numpy as np
import pytest
import torch

from oracle import batch_oracle as bo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "batch_reference_vectors.npz")
NAMES = ("s6_n4m5", "s5_n5m3_small", "s9_n8m4_odd")


def load(name):
    z = np.load(GOLD)
    S, frames, mels, N, M, L, seed = (int(v) for v in z[name + "_cfg"])
    files = [z[f"{name}_file{s}"] for s in range(S)]
    return dict(files=files, order=z[name + "_order"], perm=z[name + "_perm"], batch=z[name + "_batch"],
                frames=frames, mels=mels, N=N, M=M, L=L, seed=seed)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_dataset(name):
    g = load(name)
    spk = [g["files"][i] for i in g["order"]]
    np.random.seed(g["seed"])
    utt, clip = bo.draw_indices([a.shape[0] for a in spk], g["M"], g["frames"], g["L"])
    got = bo.assemble(spk, utt, clip, g["L"], g["perm"])
    assert got.dtype == np.float32 and np.array_equal(got, g["batch"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_bank_reproduces_reference_batches(name):
    """Same np.random / random seeds as the reference run -> the same crops, rows in the trainer's permuted
    order, float32 bits identical (mels = 40 / 8 take the 16-byte path, mels = 6 the scalar one)."""
    import speaker_embedding_ge2e_loss_b200 as pkg
    g = load(name)
    bank = pkg.SpectrogramBank(g["files"], device="cuda:0")
    np.random.seed(g["seed"])
    random.seed(g["seed"])
    batch, unperm = bank.training_batch([int(i) for i in g["order"]], g["M"], g["L"])
    torch.cuda.synchronize()
    assert batch.dtype == torch.float32 and tuple(batch.shape) == g["batch"].shape
    assert np.array_equal(batch.cpu().numpy(), g["batch"])
    perm = g["perm"]
    assert np.array_equal(unperm.cpu().numpy()[perm], np.arange(len(perm)))          # s4:184-185


@pytest.mark.gpu
@pytest.mark.parametrize("S,frames,mels,N,M,L", [(3, 180, 40, 3, 10, 160), (12, 64, 80, 12, 4, 50), (5, 33, 7, 4, 3, 9),
                                                  (2, 20, 4, 2, 2, 18), (40, 180, 40, 32, 10, 160)])
def test_gpu_assemble_bit_exact_vs_oracle(S, frames, mels, N, M, L):
    import speaker_embedding_ge2e_loss_b200 as pkg
    rng = np.random.default_rng(S * 100 + mels)
    files = [rng.standard_normal((int(rng.integers(2, 7)), frames, mels)) for _ in range(S)]
    bank = pkg.SpectrogramBank(files, device="cuda:0")
    speakers = [int(s) for s in rng.permutation(S)[:N]]
    utt = np.stack([rng.integers(0, files[s].shape[0], M) for s in speakers])
    clip = rng.integers(0, frames - L + 1, N)                     # includes the last legal start
    perm = rng.permutation(N * M)
    got = bank.assemble(speakers, utt, clip, L, perm)
    ref = bo.assemble([files[s] for s in speakers], utt, clip, L, perm)
    assert np.array_equal(got.cpu().numpy(), ref)
    out = torch.full((N * M, L, mels), float("nan"), device="cuda:0")
    bank.assemble(speakers, utt, clip, L, None, out=out)          # caller's buffer, no permutation
    assert np.array_equal(out.cpu().numpy(), bo.assemble([files[s] for s in speakers], utt, clip, L, None))


@pytest.mark.gpu
def test_gpu_bank_errors_and_from_dir(tmp_path):
    import speaker_embedding_ge2e_loss_b200 as pkg
    rng = np.random.default_rng(0)
    for s in range(3):
        np.save(tmp_path / f"sv_{s}.npy", rng.standard_normal((4, 30, 8)))
    bank = pkg.SpectrogramBank.from_dir(str(tmp_path), device="cuda:0")
    assert len(bank) == 3 and bank.frames == 30 and bank.mels == 8 and sorted(bank.names) == ["sv_0.npy", "sv_1.npy", "sv_2.npy"]
    with pytest.raises(IndexError):
        bank.assemble([0], [[4]], [0], 10)                        # utterance index out of range
    with pytest.raises(IndexError):
        bank.assemble([0], [[0]], [25], 10)                       # crop past the last frame
    with pytest.raises(ValueError):
        pkg.SpectrogramBank([rng.standard_normal((2, 30, 8)), rng.standard_normal((2, 31, 8))], device="cuda:0")
    with pytest.raises(RuntimeError):
        pkg.SpectrogramBank([rng.standard_normal((2, 30, 8))], device="cpu")
