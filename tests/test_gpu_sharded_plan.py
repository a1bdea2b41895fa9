"""ShardedGE2EPlan (the NCCL-inside-a-CUDA-graph step bench.py times at N>1) against the single-GPU
plan on the same batch.  One GPU is all the test box has, so the group has one rank: the collectives
degenerate to copies but the graph capture of C-ABI stages + NCCL, the buffer wiring and the
{loss, dw, db} all-reduce are the code the multi-GPU run uses.  Runs in a child process with a timeout so
that a communicator problem fails the test instead of hanging the suite."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent("""
    import os, sys
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from speaker_embedding_ge2e_loss_b200 import GE2EPlan, ShardedGE2EPlan
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", RANK="0", WORLD_SIZE="1")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
    dev = torch.device("cuda:0")
    def rel(a, r):
        return ((a - r).abs().max() / r.abs().max().clamp_min(1e-12)).item()
    for (N, M, D, variant, precision, tol) in [(64, 10, 256, "softmax", "tf32", 2e-3),
                                                (64, 10, 256, "contrast", "tf32", 2e-3),
                                                (48, 8, 128, "softmax", "fp32", 1e-5),
                                                (512, 4, 256, "softmax", "tf32", 2e-3)]:    # tcgen05 step kernel
        g = torch.Generator().manual_seed(N + M)
        Es = [torch.nn.functional.normalize(torch.randn(N, M, D, generator=g), dim=-1).to(dev) for _ in range(2)]
        w = torch.tensor(10.0, device=dev); b = torch.tensor(-5.0, device=dev)
        ref = GE2EPlan(N, M, D, variant, precision, device=dev)
        sp = ShardedGE2EPlan(N, N, 0, M, D, variant, precision, device=dev)
        graph = sp.capture(Es, w, b, steps=2)       # step 0 on Es[0], step 1 on Es[1]
        graph.replay(); torch.cuda.synchronize()
        ref.step(Es[1], w, b); torch.cuda.synchronize()
        errs = dict(loss=rel(sp.loss, ref.loss), dE=rel(sp.dE, ref.dE.view_as(sp.dE)), dw=rel(sp.dw, ref.dw),
                    db=rel(sp.db, ref.db))
        assert all(e <= tol for e in errs.values()), (N, M, D, variant, precision, errs)
        sp.step(Es[0], w, b); ref.step(Es[0], w, b); torch.cuda.synchronize()     # eager step, same buffers
        assert rel(sp.dE, ref.dE.view_as(sp.dE)) <= tol and rel(sp.loss, ref.loss) <= tol
        del graph
    # host-fed sharded plan: pinned shards in, global {loss, dw, db} out, more batches than slots
    from speaker_embedding_ge2e_loss_b200 import ShardedGE2EHostFeed
    N, M, D = 64, 10, 256
    w = torch.tensor(10.0, device=dev); b = torch.tensor(-5.0, device=dev)
    g = torch.Generator().manual_seed(7)
    hosts = [torch.nn.functional.normalize(torch.randn(N, M, D, generator=g), dim=-1).pin_memory() for _ in range(5)]
    feed = ShardedGE2EHostFeed(N, N, 0, M, D, w, b, "softmax", "tf32", device=dev)
    ref = GE2EPlan(N, M, D, "softmax", "tf32", device=dev)
    prev, got, dEs = None, [], []
    for h in hosts:
        t = feed.submit(h)
        if prev is not None:
            got.append(feed.result(prev))
        feed.done_event(t).synchronize()
        dEs.append(feed.dE(t).clone())
        prev = t
    got.append(feed.result(prev))
    for k, h in enumerate(hosts):
        ref.step(h.to(dev), w, b); torch.cuda.synchronize()
        assert abs(got[k][0] - ref.loss.item()) <= 2e-3 * abs(ref.loss.item()), (k, got[k], ref.loss.item())
        assert abs(got[k][1] - ref.dw.item()) <= 2e-3 * max(1.0, abs(ref.dw.item()))
        assert rel(dEs[k], ref.dE.view_as(dEs[k])) <= 2e-3
    print("SHARDED_PLAN_OK", flush=True)
    os._exit(0)
""") % ROOT


@pytest.mark.gpu
def test_sharded_plan_world1_matches_single_gpu_plan():
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=300)
    assert "SHARDED_PLAN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
