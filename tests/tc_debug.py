"""Stage-by-stage comparison of the tcgen05 (TF32) kernels against the SIMT fp32 kernels on the
same device buffers.  Debug aid, not a pytest module:  python tests/tc_debug.py [N M D [variant]]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ge2e_oracle as orc  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import _lib, lib, ops  # noqa: E402


def summarize(name, got, want, rows=None):
    got = got.double().cpu().numpy()
    want = want.double().cpu().numpy()
    err = np.abs(got - want)
    rel = np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)
    print(f"  {name:10s} rel_l2={rel:.3e} max_abs={err.max():.3e} |want|max={np.abs(want).max():.3e} "
          f"nan={np.isnan(got).sum()}")
    if rel > 5e-3 or np.isnan(got).any():
        bad = np.argwhere((err > 1e-2 * max(np.abs(want).max(), 1e-30)) | np.isnan(got))
        print(f"    bad elements: {len(bad)} of {got.size}; first {bad[:8].tolist()}")
        if got.ndim == 2:
            br = np.unique(bad[:, 0])
            bc = np.unique(bad[:, 1])
            print(f"    bad rows: n={len(br)} head={br[:12].tolist()} tail={br[-4:].tolist()}")
            print(f"    bad cols: n={len(bc)} head={bc[:12].tolist()} tail={bc[-4:].tolist()}")
            r0 = bad[0][0]
            print(f"    row {r0} got  {got[r0, :8]}")
            print(f"    row {r0} want {want[r0, :8]}")
    return rel


def main():
    N, M, D = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 10, 256)
    variant = _lib.VARIANTS[sys.argv[4]] if len(sys.argv) >= 5 else _lib.SOFTMAX
    dev = torch.device("cuda:0")
    print(f"N={N} M={M} D={D} variant={variant} path(tf32)={lib().ge2e_b200_path(N, N, M, D, variant, _lib.TF32)}")
    E = torch.tensor(orc.make_embeddings(N, M, D, seed=1, kind="clustered"), device=dev)
    w = torch.tensor(10.0, device=dev)
    b = torch.tensor(-5.0, device=dev)
    g = torch.tensor(1.0, device=dev)
    eps = 1e-6
    c_hat = torch.empty((N, D), device=dev)
    e_hat, cos_diag, accum = ops.prep(E, c_hat, _lib.TF32)      # tf32-rounded operands for both paths
    out = {}
    for tag, prec in (("simt", _lib.FP32), ("tc", _lib.TF32)):
        accum.zero_()
        rs, ks, aux, per, _ = ops.fwd_rows(e_hat, c_hat, cos_diag, N, N, 0, M, D, w, b, eps, variant, prec, accum,
                                           per_row=True)
        torch.cuda.synchronize()
        out[tag] = dict(rs=rs.clone(), per=per.clone(), loss=accum[0].item(), ks=ks.clone(), aux=aux.clone())
        print(f"  [{tag}] nan counts: row_stat={torch.isnan(rs).sum().item()} per={torch.isnan(per).sum().item()} "
              f"aux={torch.isnan(aux).sum().item()}")
    print("forward:")
    print(f"  loss simt={out['simt']['loss']:.6f} tc={out['tc']['loss']:.6f}")
    summarize("row_stat", out["tc"]["rs"], out["simt"]["rs"])
    summarize("per_row", out["tc"]["per"], out["simt"]["per"])
    summarize("row_aux", out["tc"]["aux"], out["simt"]["aux"])
    if variant == _lib.CONTRAST:
        print("  kstar mismatches:", (out["tc"]["ks"] != out["simt"]["ks"]).sum().item())
        return
    rs = out["simt"]["rs"]
    bw = {}
    for tag, prec in (("simt", _lib.FP32), ("tc", _lib.TF32)):
        dE_hat, dC, dwdb = ops.bwd_rows(e_hat, c_hat, cos_diag, rs, out["simt"]["ks"], out["simt"]["aux"], N, N, 0,
                                        M, D, w, b, eps, variant, prec, g)
        torch.cuda.synchronize()
        bw[tag] = (dE_hat.clone(), dC.clone(), dwdb.clone())
    print("backward:")
    summarize("dE_hat", bw["tc"][0], bw["simt"][0])
    summarize("dC_hat", bw["tc"][1], bw["simt"][1])
    print(f"  dwdb simt={bw['simt'][2].tolist()} tc={bw['tc'][2].tolist()}")


if __name__ == "__main__":
    main()
