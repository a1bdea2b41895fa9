"""The UNMODIFIED reference trainer and evaluation code (baseline/_ref, installed by
scripts/install_reference.py) running with this package's GE2ELoss substituted for its own.

  * ``TrainEmbedModel(hp).train_model(...)`` (s4_train_embed_model.py:17-45, :137-272) is executed twice on
    the same seeded synthetic spectrogram batches -- once with the reference's ``GE2ELoss``
    (s3_loss_function_GE2E.py, plain torch on the GPU), once with ours injected the way INTEGRATION.md
    describes (``s4.GE2ELoss = ours``) -- and the per-epoch training losses, the test losses and the learned
    ``w`` / ``b`` must track each other.  Only ``get_train_test_data_loader`` (s1_dataset_loader.py:82-108:
    needs .npy corpora on disk and is broken on torch >= 2) is replaced, by a list of tensors.
  * ``calculate_ERR`` (s5_eval_model.py:16-100) with the reference's static helpers replaced by ours prints
    the same EER line as the reference's own run, and ``evaluate_eer`` returns the same numbers.
"""
import io
import os
import random
import re
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def ref():
    if not os.path.isdir(os.path.join(REF, "embedding_model_GE2E")):
        pytest.skip("baseline/_ref is not installed (python scripts/install_reference.py)")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # s5 imports matplotlib at module level for its t-SNE plots (not installed here, not on this path)
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import embedding_model_GE2E.s3_loss_function_GE2E as s3
    import embedding_model_GE2E.s4_train_embed_model as s4
    import embedding_model_GE2E.s5_eval_model as s5
    from utils.dict_to_dot import GetDictWithDotNotation
    return types.SimpleNamespace(s3=s3, s4=s4, s5=s5, hp_of=GetDictWithDotNotation)


def make_hp(ref, tmp, N, M, tN, tM, emb, precision=None):
    general = {"device": torch.device("cuda:0"), "small_err": 1e-6, "project_root": str(tmp)}
    if precision is not None:
        general["ge2e_precision"] = precision
    return ref.hp_of({
        "general": general,
        "audio": {"mel_n_channels": 40},
        "m_ge2e": {"restore_existing_model": False, "model_path": "none.pth", "model_hidden_size": 48,
                   "model_num_layers": 2, "model_embedding_size": emb, "lr": 0.01, "training_N": N, "training_M": M,
                   "test_N": tN, "test_M": tM, "training_epochs": 4, "checkpoint_dir": "ckpt",
                   "checkpoint_interval": 1000, "save_best_weights": False, "min_test_loss": 0.0,
                   "tt_data": {"train_spects_path": "x", "test_spects_path": "x", "min_train_utter_len": 24,
                               "min_test_utter_len": 24}},
    })


def run_trainer(ref, loss_cls, hp, batches, test_batches, seed=0):
    torch.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.backends.cudnn.deterministic = True
    ref.s4.GE2ELoss = loss_cls                                         # what s4:33 resolves at call time
    ref.s4.get_train_test_data_loader = lambda _hp: (batches, test_batches)
    trainer = ref.s4.TrainEmbedModel(hp)
    with redirect_stdout(io.StringIO()):
        _, train_losses, test_losses = trainer.train_model(lr_reduce=1000, epoch_print=2, dot_print=1)
    return (np.asarray(train_losses, dtype=np.float64), np.asarray(test_losses, dtype=np.float64),
            trainer.ge2e_loss.w.item(), trainer.ge2e_loss.b.item())


@pytest.mark.parametrize("N,M,emb,precision,tol", [(8, 5, 64, None, 2e-4), (64, 10, 256, None, 2e-4),
                                                   (256, 2, 64, "tf32", 5e-3)])
def test_reference_trainer_runs_unchanged_with_the_drop_in(ref, tmp_path, N, M, emb, precision, tol):
    import speaker_embedding_ge2e_loss_b200 as pkg
    g = torch.Generator().manual_seed(1)
    # the dataset yields float64 [speakers, utterances, frames, mels] (s1_dataset_loader.py:59-79)
    batches = [torch.randn(N, M, 24, 40, generator=g, dtype=torch.float64) for _ in range(2)]
    test_batches = [torch.randn(4, 4, 24, 40, generator=g, dtype=torch.float64)]
    hp_ref = make_hp(ref, tmp_path / "ref", N, M, 4, 4, emb)
    hp_new = make_hp(ref, tmp_path / "new", N, M, 4, 4, emb, precision)
    want = run_trainer(ref, ref.s3.GE2ELoss, hp_ref, batches, test_batches)
    got = run_trainer(ref, pkg.GE2ELoss, hp_new, batches, test_batches)
    assert np.all(np.isfinite(want[0])) and len(want[0]) == 4 and len(want[1]) == 2
    np.testing.assert_allclose(got[0], want[0], rtol=tol)             # training loss per epoch
    np.testing.assert_allclose(got[1], want[1], rtol=tol)             # test loss at epochs 2 and 4
    assert abs(got[2] - want[2]) <= tol * 10 and abs(got[3] - want[3]) <= tol * 10, (got[2:], want[2:])
    # s4:35-42 puts w, b in the optimiser (lr 0.01, their gradient clipped to norm 1 at s4:202): they moved
    assert abs(want[2] - 10.0) > 1e-5 and abs(got[2] - 10.0) > 1e-5, (want[2], got[2])
    if precision == "tf32":
        crit = pkg.GE2ELoss(hp_new)
        assert crit.precision == "tf32" and crit.path_for(N, M, emb) == 1, "expected the tcgen05 path"


def test_reference_eer_code_runs_unchanged_with_the_drop_in(ref, tmp_path):
    import speaker_embedding_ge2e_loss_b200 as pkg
    N, M, emb = 4, 16, 64
    hp = make_hp(ref, tmp_path, 8, 5, N, M, emb)
    torch.manual_seed(3)
    model = ref.s4.ModelGE2ELossSpeachEmbed(hp).to(hp.general.device).eval()
    g = torch.Generator().manual_seed(5)
    # a speaker-dependent offset makes the similarity matrix non-trivial around the 0.5 .. 0.99 thresholds
    mel = (torch.randn(N, 1, 24, 40, generator=g) * 2 + 0.3 * torch.randn(N, M, 24, 40, generator=g)).double()
    ref.s5.get_train_test_data_loader = lambda hp: (None, [mel])

    def run(loss_cls):
        ref.s5.GE2ELoss = loss_cls
        buf = io.StringIO()
        with redirect_stdout(buf), torch.no_grad():
            ref.s5.calculate_ERR(model, hp, N=N, M=M)
        return re.findall(r"EER : .*", buf.getvalue())[-1]

    want = run(ref.s3.GE2ELoss)
    got = run(pkg.GE2ELoss)
    assert got == want, (got, want)
    with torch.no_grad():
        emb_t = model(mel.reshape(N * M, 24, 40).to(hp.general.device)).reshape(N, M, emb)
    r = pkg.evaluate_eer(emb_t)
    assert "EER : %0.2f (thres:%0.2f, FAR:%0.2f, FRR:%0.2f)" % (r.EER, r.thres, r.FAR, r.FRR) == want
