# one forward + backward of the model tail at config 3's shape: the launch list must hold no library GEMM
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import speaker_embedding_ge2e_loss_b200 as pkg
dev = torch.device("cuda:0")
U, frames, H, D = 10240, 2, 768, 256
tail = pkg.ProjectionL2Norm(H, D).to(dev)
out = torch.randn(U, frames, H, device=dev, requires_grad=True)
dE = torch.randn(U, D, device=dev)
for _ in range(2):
    out.grad = None
    (tail(out) * dE).sum().backward()
torch.cuda.synchronize()
print("ok", float(out.grad[:, -1].abs().max()), float(tail.projection.weight.grad.abs().max()))
