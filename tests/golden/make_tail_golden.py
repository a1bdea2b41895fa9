"""Generate tests/golden/tail_reference_vectors.npz from the REAL reference model class
(ModelGE2ELossSpeachEmbed: LSTM stack + projection + L2 normalise, s2:7-35) run on CPU in this
container: the LSTM's last-frame activations (captured by a forward hook), the projection's weight and
bias, the embeddings the module returns, and torch-autograd gradients wrt those three for a seeded
upstream gradient.  Run only where /root/reference is mounted:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_tail_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from embedding_model_GE2E.s2_model_GE2E_loss_speach_embed import ModelGE2ELossSpeachEmbed  # noqa: E402
from utils.dict_to_dot import GetDictWithDotNotation  # noqa: E402


def run(U, frames, mels, hidden, layers, emb, seed, dtype):
    torch.manual_seed(seed)
    hp = GetDictWithDotNotation({"audio": {"mel_n_channels": mels},
                                 "m_ge2e": {"model_hidden_size": hidden, "model_num_layers": layers,
                                            "model_embedding_size": emb}})
    model = ModelGE2ELossSpeachEmbed(hp).to(dtype)
    grabbed = {}

    def hook(_m, _inp, out):
        out[0].retain_grad()
        grabbed["lstm_out"] = out[0]

    model.LSTM_stack.register_forward_hook(hook)
    x = torch.randn(U, frames, mels, dtype=dtype)
    if dtype == torch.float64:
        # s2:28 / :31 call .float(): run the same three lines in float64 for a "truth" copy
        out, _ = model.LSTM_stack(x)
        last = out[:, out.size(1) - 1]
        y = model.projection(last)
        E = y / torch.norm(y, dim=1).unsqueeze(1)
    else:
        E = model(x)
    dE = torch.randn(E.shape, dtype=dtype)
    (E * dE).sum().backward()
    lstm_out = grabbed["lstm_out"]
    return dict(x_last=lstm_out[:, -1].detach().numpy(), W=model.projection.weight.detach().numpy(),
                bias=model.projection.bias.detach().numpy(), E=E.detach().numpy(), dE=dE.numpy(),
                dX=lstm_out.grad[:, -1].numpy(), dW=model.projection.weight.grad.numpy(),
                dbias=model.projection.bias.grad.numpy())


def main():
    out = {}
    cases = [("u40_h768_d64", 40, 12, 40, 768, 1, 64, 1), ("u130_h96_d64", 130, 6, 20, 96, 2, 64, 2),
             ("u7_h64_d128", 7, 5, 16, 64, 1, 128, 3), ("u200_h128_d256", 200, 4, 24, 128, 1, 256, 4)]
    for name, U, frames, mels, hidden, layers, emb, seed in cases:
        # float32 = what a user of the reference sees; a float64 run of the same module for the smallest case
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64))[:2 if U < 10 else 1]:
            r = run(U, frames, mels, hidden, layers, emb, seed, dt)
            for k, v in r.items():
                out[f"{name}_{tag}_{k}"] = v
        print(name, r["x_last"].shape, r["W"].shape, float(np.abs(r["E"]).max()))
    np.savez_compressed(os.path.join(HERE, "tail_reference_vectors.npz"), **out)


if __name__ == "__main__":
    main()
