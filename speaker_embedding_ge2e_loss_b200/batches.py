"""Training batches assembled on the GPU from a device-resident spectrogram bank.

Reference: the dataset (embedding_model_GE2E/s1_dataset_loader.py:52-79) re-reads a speaker's
``float64[utts, frames, mels]`` file for every item, picks ``M`` random utterances and one random crop
of ``L`` frames; the DataLoader stacks ``N`` speakers, and the trainer (s4_train_embed_model.py:170-186)
copies the float64 batch to the device, reshapes it to [N*M, L, mels] and permutes its rows before the
model's ``x.float()`` (s2:28).  Here every speaker's array is converted to float32 ONCE and kept in HBM
(the corpus is a few GB; the GPU has 180 GB), the host only draws the indices -- with the reference's
own two ``np.random.randint`` calls per speaker, in its order, so a seeded run picks the same crops --
and ONE kernel (``ge2e_b200_gather_spans``) writes the model's float32 input batch, row permutation
included.  float64 -> float32 is the same round-to-nearest the reference applies, so the batch is
bit-identical to the reference's.  CUDA (sm_100a) only; no host fallback.

    bank = SpectrogramBank.from_dir(train_specs_path, device="cuda")     # once
    spk = bank.speaker_order[k * N:(k + 1) * N]
    batch, unperm = bank.training_batch(spk, M=10, crop_len=160)          # [N*M, 160, mels], int32 [N*M]
    E = model(batch)
    loss = crit(E, unperm=unperm, speakers=N)                             # s4:189-192 folded into the loss
"""
from __future__ import annotations

import os
import random
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops


def span_offsets(first, frames: int, mels: int, speakers, utter_idx, clip, perm=None) -> np.ndarray:
    """Element offset, in a bank laid out as [all utterances, frames, mels], of every output row's crop:
    row (n, m) = utterance ``first[speakers[n]] + utter_idx[n, m]`` starting at frame ``clip[n]``; rows
    reordered by ``perm`` (s4:186 ``mel_db_batch[perm]``).  Host only (pure index arithmetic)."""
    first = np.asarray(first, dtype=np.int64)
    spk = np.asarray(speakers, dtype=np.int64)
    off = ((first[spk][:, None] + np.asarray(utter_idx, dtype=np.int64)) * frames + np.asarray(clip, dtype=np.int64)[:, None]) * mels
    off = off.reshape(-1)
    if perm is not None:
        off = off[np.asarray(perm, dtype=np.int64)]
    return np.ascontiguousarray(off)


class SpectrogramBank:
    def __init__(self, arrays: Sequence[np.ndarray], device="cuda", names: Optional[List[str]] = None):
        if len(arrays) == 0:
            raise ValueError("SpectrogramBank needs at least one speaker array")
        frames, mels = arrays[0].shape[1], arrays[0].shape[2]
        for a in arrays:
            if a.ndim != 3 or a.shape[1] != frames or a.shape[2] != mels:
                raise ValueError("every speaker array must be [utts, frames, mels] with the same frames and mels")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SpectrogramBank lives in GPU memory (sm_100a); there is no host fallback")
        self.frames, self.mels = int(frames), int(mels)
        self.utts = np.asarray([a.shape[0] for a in arrays], dtype=np.int64)
        self.first = np.concatenate([[0], np.cumsum(self.utts)]).astype(np.int64)       # first utterance of a speaker
        self.names = names
        total = int(self.first[-1])
        self.data = torch.empty((total, self.frames, self.mels), dtype=torch.float32, device=self.device)
        for s, a in enumerate(arrays):                                                  # s2:28 `.float()`, done once
            self.data[int(self.first[s]):int(self.first[s + 1])].copy_(torch.from_numpy(np.asarray(a, dtype=np.float32)))
        self.speaker_order = list(range(len(arrays)))

    @classmethod
    def from_dir(cls, data_path: str, device="cuda", shuffle: bool = False):
        """The reference's file discovery (s1:21-29: the files of the top-level directory, in os.walk order);
        ``shuffle=True`` applies its ``random.shuffle`` of the speaker list (s1:40)."""
        files = next(iter(os.walk(data_path)))[2]
        bank = cls([np.load(os.path.join(data_path, f)) for f in files], device=device, names=list(files))
        if shuffle:
            random.shuffle(bank.speaker_order)
        return bank

    def __len__(self):
        return len(self.utts)

    def draw(self, speakers: Sequence[int], M: int, crop_len: int, rng=np.random):
        """The dataset's two random draws per speaker (s1:66, :72), in its order: (utter_idx[N, M], clip[N])."""
        utt, clip = [], []
        for s in speakers:
            utt.append(rng.randint(0, int(self.utts[s]), M))
            clip.append(rng.randint(0, self.frames - crop_len - 1))
        return np.asarray(utt, dtype=np.int64).reshape(len(speakers), M), np.asarray(clip, dtype=np.int64)

    def assemble(self, speakers: Sequence[int], utter_idx, clip, crop_len: int, perm=None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """float32 [N*M, crop_len, mels]: row r is utterance ``perm[r]`` of the [N, M] grid (s4:176-186)."""
        spk = np.asarray(list(speakers), dtype=np.int64)
        utter_idx = np.asarray(utter_idx, dtype=np.int64)
        clip = np.asarray(clip, dtype=np.int64)
        N, M = utter_idx.shape
        if spk.shape != (N,) or clip.shape != (N,):
            raise ValueError("speakers [N], utter_idx [N, M] and clip [N] do not agree")
        if (spk < 0).any() or (spk >= len(self)).any() or (utter_idx < 0).any() or (utter_idx >= self.utts[spk][:, None]).any():
            raise IndexError("speaker or utterance index out of range")
        if crop_len < 1 or (clip < 0).any() or (clip + crop_len > self.frames).any():
            raise IndexError("crop outside the utterance")
        off = span_offsets(self.first, self.frames, self.mels, spk, utter_idx, clip, perm)
        rows, span = N * M, crop_len * self.mels
        if out is None:
            out = torch.empty((rows, crop_len, self.mels), dtype=torch.float32, device=self.device)
        elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == rows * span):
            raise ValueError("out must be a contiguous float32 CUDA tensor of N*M*crop_len*mels values")
        with torch.cuda.device(self.device):
            off_dev = torch.from_numpy(np.ascontiguousarray(off)).to(self.device, non_blocking=False)
            _lib.check(_lib.lib().ge2e_b200_gather_spans(self.data.data_ptr(), off_dev.data_ptr(), rows, span,
                                                         1 if self.mels % 4 == 0 else 0, out.data_ptr(),
                                                         ops._stream()), "ge2e_b200_gather_spans")
        return out

    def training_batch(self, speakers: Sequence[int], M: int, crop_len: int, rng=np.random, py_random=random,
                       out: Optional[torch.Tensor] = None):
        """One trainer iteration's input (s1:59-77 + s4:176-186): draws, row permutation
        ``perm = random.sample(range(N*M), N*M)``, gather.  Returns (batch [N*M, crop_len, mels],
        unperm int32 [N*M] on the device) -- pass ``unperm`` to ``GE2ELoss.forward(..., unperm=, speakers=N)``."""
        utter_idx, clip = self.draw(speakers, M, crop_len, rng)
        total = len(speakers) * M
        perm = py_random.sample(range(0, total), total)                  # s4:179
        unperm = np.empty(total, dtype=np.int32)
        unperm[np.asarray(perm)] = np.arange(total, dtype=np.int32)      # s4:184-185
        batch = self.assemble(speakers, utter_idx, clip, crop_len, perm, out)
        return batch, torch.from_numpy(unperm).to(self.device)
