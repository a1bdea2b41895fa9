import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import speaker_embedding_ge2e_loss_b200 as pkg
from oracle import ge2e_oracle as orc
from oracle import ge2e_oracle_torch as orct
DEV='cuda:0'
def trel(a,b): return ((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30)).item()
def run_module(E, w, b, precision, g=None):
    crit = pkg.GE2ELoss(None, device=E.device, w=w, b=b, variant="softmax", precision=precision)
    Eg = E.clone().requires_grad_(True)
    loss = crit(Eg)
    (loss if g is None else loss * g).backward()
    torch.cuda.synchronize()
    return dict(loss=loss.item(), dE=Eg.grad, dw=crit.w.grad.item(), db=crit.b.grad.item())
for (N,M,D,w,b,g,kind) in [(384,4,128,30.0,-10.0,-1.5,'clustered'),(384,4,128,10.0,-5.0,1.0,'clustered'),(1024,10,256,10.0,-5.0,1.0,'clustered'),(1024,10,256,30.0,-10.0,1.0,'clustered'),(1024,10,256,10.0,-5.0,1.0,'random')]:
    E = torch.tensor(orc.make_embeddings(N, M, D, seed=21, kind=kind), device=DEV)
    ref = orct.forward_backward(E, w, b, 1e-6, "softmax", g=g)
    for prec in ("fp32", "fp32_simt", "tf32"):
        got = run_module(E, w, b, prec, g=g)
        # per-row relative error distribution
        d = (got["dE"].double()-ref["dE"]).reshape(N*M,-1).norm(dim=1)
        rn = ref["dE"].reshape(N*M,-1).norm(dim=1)
        print(N,M,D,w,kind,prec, "dE_rel %.3e"%trel(got["dE"],ref["dE"]), "loss_rel %.2e"%(abs(got["loss"]-ref["loss"])/abs(ref["loss"])),
              "dw_rel %.2e"%(abs(got["dw"]-ref["dw"])/max(1,abs(ref["dw"]))), "row-rel median %.2e max %.2e"%((d/rn).median().item(), (d/rn).max().item()))
