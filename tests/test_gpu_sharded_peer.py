"""Speaker-sharded step over peer memory on 2 real GPUs (skipped on a single-GPU box): the plan whose exchange
steps are done by the kernels themselves (ge2e_b200_peer_publish + ge2e_b200_step_rows_peers between two
symmetric-memory barriers) must reproduce the single-GPU plan on the concatenated batch, eagerly and as a
captured graph, and agree with the NCCL (all-gather / reduce-scatter) plan."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from speaker_embedding_ge2e_loss_b200 import GE2EPlan, ShardedGE2EPlan
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    N, M, D = 1024, 6, 256
    nl, off = N // world, rank * (N // world)
    g = torch.Generator().manual_seed(3)
    E_all = [torch.nn.functional.normalize(torch.randn(N, M, D, generator=g), dim=-1).to(dev) for _ in range(2)]
    w = torch.tensor(10.0, device=dev); b = torch.tensor(-5.0, device=dev)
    ref = GE2EPlan(N, M, D, "softmax", "tf32", device=dev)
    peer = ShardedGE2EPlan(nl, N, off, M, D, "softmax", "tf32", device=dev, peer_memory=True)
    nccl = ShardedGE2EPlan(nl, N, off, M, D, "softmax", "tf32", device=dev, peer_memory=False)
    assert peer.peer and not nccl.peer
    def rel(a, r):
        return ((a.double() - r.double()).norm() / r.double().norm().clamp_min(1e-30)).item()
    for k in (0, 1, 0):
        shard = E_all[k][off:off + nl].contiguous()
        ref.step(E_all[k], w, b); peer.step(shard, w, b); nccl.step(shard, w, b)
        torch.cuda.synchronize()
        for p in (peer, nccl):
            assert rel(p.dE, ref.dE[off:off + nl]) <= 2e-5, (k, rel(p.dE, ref.dE[off:off + nl]))
            assert abs(p.loss.item() - ref.loss.item()) <= 1e-5 * abs(ref.loss.item())
            assert abs(p.dw.item() - ref.dw.item()) <= 1e-4 * max(1.0, abs(ref.dw.item()))
    # the default precision (fp32-class on the tensor cores: centroid rows carry their hi / lo fp16 planes
    # through the publish / the all-gather)
    ref32 = GE2EPlan(N, M, D, "softmax", "fp32", device=dev)
    peer32 = ShardedGE2EPlan(nl, N, off, M, D, "softmax", "fp32", device=dev, peer_memory=True)
    nccl32 = ShardedGE2EPlan(nl, N, off, M, D, "softmax", "fp32", device=dev, peer_memory=False)
    assert ref32.precision == 2 and peer32.precision == 2 and peer32.peer and not nccl32.peer
    for k in (1, 0):
        shard = E_all[k][off:off + nl].contiguous()
        ref32.step(E_all[k], w, b); peer32.step(shard, w, b); nccl32.step(shard, w, b)
        torch.cuda.synchronize()
        for p in (peer32, nccl32):
            assert rel(p.dE, ref32.dE[off:off + nl]) <= 5e-6, (k, rel(p.dE, ref32.dE[off:off + nl]))
            assert abs(p.loss.item() - ref32.loss.item()) <= 2e-6 * abs(ref32.loss.item())
            assert abs(p.dw.item() - ref32.dw.item()) <= 1e-5 * max(1.0, abs(ref32.dw.item()))
    shards = [E_all[k][off:off + nl].contiguous() for k in range(2)]
    graph = peer.capture(shards, w, b, steps=4)          # steps 0..3 on shards 0, 1, 0, 1
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    ref.step(E_all[1], w, b); torch.cuda.synchronize()
    assert rel(peer.dE, ref.dE[off:off + nl]) <= 2e-5 and abs(peer.loss.item() - ref.loss.item()) <= 1e-5 * abs(ref.loss.item())
    dist.barrier()
    if rank == 0:
        print("SHARDED_PEER_OK", flush=True)
    del graph
    torch.cuda.synchronize()
    os._exit(0)
""") % ROOT


@pytest.mark.gpu
def test_peer_memory_sharded_plan_on_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "peer_child.py"
    script.write_text(CHILD)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29588", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert "SHARDED_PEER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
