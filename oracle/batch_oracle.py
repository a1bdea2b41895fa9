"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's batch assembly.

Follows
  * /root/reference/embedding_model_GE2E/s1_dataset_loader.py:52-79 (``__getitem__``): load the
    speaker's ``float64[utts, frames, mels]`` array, ``utter_idx = np.random.randint(0, utts, M)``,
    gather those utterances, ``random_clip = np.random.randint(0, frames - min_utter_len - 1)``, crop
    ``[:, clip:clip + min_utter_len, :]``;
  * the DataLoader's default collate: stack N speakers -> float64 [N, M, L, mels];
  * /root/reference/embedding_model_GE2E/s4_train_embed_model.py:170-186: ``.to(device)``, reshape to
    [N*M, L, mels], ``mel_db_batch[perm]``; and s2:28 ``x.float()`` (float64 -> float32, round to nearest)
    as the model's first operation.
Pinned against the reference's own dataset class run on seeded files:
tests/golden/make_batch_golden.py -> tests/golden/batch_reference_vectors.npz.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np


def draw_indices(utts_per_speaker, M, frames, crop_len, rng=np.random):
    """The two np.random.randint calls of ``__getitem__`` (s1:66, :72) per speaker, in its order.
    Returns (utter_idx[N, M], clip[N])."""
    utt, clip = [], []
    for n_utts in utts_per_speaker:
        utt.append(rng.randint(0, n_utts, M))
        clip.append(rng.randint(0, frames - crop_len - 1))
    return np.asarray(utt, dtype=np.int64), np.asarray(clip, dtype=np.int64)


def get_item(spr_utters, utter_idx, random_clip, crop_len):
    """s1:69-77 for one speaker."""
    return spr_utters[utter_idx, :, :][:, random_clip:random_clip + crop_len, :]


def assemble(speakers, utter_idx, clip, crop_len, perm=None):
    """float32 [N*M, crop_len, mels] as the model sees it (s4:170-186 + s2:28): collate, reshape, permute, .float()."""
    batch = np.stack([get_item(s, utter_idx[i], int(clip[i]), crop_len) for i, s in enumerate(speakers)])
    N, M = batch.shape[:2]
    flat = batch.reshape(N * M, batch.shape[2], batch.shape[3])
    if perm is not None:
        flat = flat[np.asarray(perm)]
    return flat.astype(np.float32)
