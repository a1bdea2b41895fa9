import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import speaker_embedding_ge2e_loss_b200 as pkg
from speaker_embedding_ge2e_loss_b200 import ops
from oracle import ge2e_oracle as orc
DEV='cuda:0'
def lse64(E, w, b, eps=1e-6):
    E = E.double(); N, M, D = E.shape
    C = E.mean(1)
    Ssum = E.sum(1, keepdim=True)
    Uc = (Ssum - E) / (M - 1)
    cosf = torch.nn.functional.cosine_similarity
    En = E / E.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    Cn = C / C.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    cos = En.reshape(N*M, D) @ Cn.t()
    cd = cosf(E, Uc, dim=-1).reshape(N*M)
    idx = torch.arange(N*M, device=E.device)
    cos[idx, idx // M] = cd
    S = w * (cos + eps) + b
    ex = torch.exp(S)
    q = (ex.sum(1) - ex[idx, idx // M] + eps) / (ex.sum(1) + eps)
    return (torch.log(ex.sum(1) + eps), q), cos
for (N,M,D,w,b) in [(384,4,128,30.0,-10.0),(384,4,256,30.0,-10.0),(1024,10,256,30.0,-10.0),(1024,10,256,10.0,-5.0)]:
    E = torch.tensor(orc.make_embeddings(N, M, D, seed=21, kind='clustered'), device=DEV)
    (ref, q64), cos = lse64(E, w, b)
    wt, bt = torch.tensor(w, device=DEV), torch.tensor(b, device=DEV)
    offd = cos.clone(); idx = torch.arange(N*M, device=DEV); offd[idx, idx//M] = float('nan')
    print("offdiag cos mean %.3f max %.3f" % (torch.nanmean(offd).item(), offd.nan_to_num(-1).max().item()))
    for prec in (2, 0, 1):
        c_hat = torch.empty((N, D), device=DEV)
        e_hat, cos_diag, accum = ops.prep(E, c_hat, prec)
        row_stat, _, row_aux, *_ = ops.fwd_rows(e_hat, c_hat, cos_diag, N, N, 0, M, D, wt, bt, 1e-6, 0, prec, accum)
        torch.cuda.synchronize()
        err = (row_stat.double() - ref)
        qe = (row_aux.double() - q64) / q64
        print("   q rel err mean %.3e std %.3e median|.| %.2e" % (qe.mean().item(), qe.std().item(), qe.abs().median().item()))
        print(N,M,D,w,"prec",prec,"lse err mean %.3e std %.3e  (ulp(lse)=%.1e)"%(err.mean().item(), err.std().item(), 2.0**(torch.log2(ref.abs().mean()).floor().item()-23)))
