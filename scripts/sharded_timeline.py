"""Per-stage timeline of one speaker-sharded step (BASELINE config 4 by default), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port 29511 \
        scripts/sharded_timeline.py [--shape 8192,16,256] [--iters 30]

The stages of ShardedGEPlan.step (prep | all-gather c_hat | step_rows | reduce-scatter dC_hat | all-reduce
{loss, dw, db} | bwd_finalize) are issued eagerly on one stream with a CUDA event between every two of them;
the table is the median over the iterations of each stage's duration, max over ranks, plus the graph-replay
time of the whole step for comparison (eager launch gaps are NOT in the graph).  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speaker_embedding_ge2e_loss_b200 import ShardedGE2EPlan, lib  # noqa: E402
from speaker_embedding_ge2e_loss_b200._lib import check  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="8192,16,256")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--no-peer", action="store_true", help="NCCL all-gather / reduce-scatter instead of peer memory")
    a = ap.parse_args()
    N, M, D = (int(v) for v in a.shape.split(","))
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nl = N // world
    off = rank * nl
    g = torch.Generator().manual_seed(rank)
    E = torch.nn.functional.normalize(torch.randn(nl, M, D, generator=g), dim=-1).to(dev)
    w, b = torch.tensor(10.0, device=dev), torch.tensor(-5.0, device=dev)
    p = ShardedGE2EPlan(nl, N, off, M, D, "softmax", "tf32", device=dev, peer_memory=False if a.no_peer else "auto")
    h = lib()
    ws = p._ws.data_ptr() if p._ws_bytes else None
    names = (["prep", "peer_publish", "barrier", "step_rows_peers", "scalars+barrier2", "sum_scalars", "bwd_finalize"] if p.peer
             else ["prep", "all_gather", "step_rows", "reduce_scatter", "all_reduce", "bwd_finalize"])

    def stages_peer():
        s = torch.cuda.current_stream(dev).cuda_stream
        yield lambda: check(h.ge2e_b200_prep(E.data_ptr(), nl, M, D, p.precision, p.e_hat.data_ptr(),
                                             p.c_hat_mine.data_ptr(), p.cos_diag.data_ptr(), p.red.data_ptr(), s), "prep")
        yield lambda: check(h.ge2e_b200_peer_publish(p.c_hat_mine.data_ptr(), p._peer_slices, p._n_peers, p._mcast, nl * D,
                                                     p.dC_local.data_ptr(), nl * D, s), "publish")
        yield lambda: p._hc.barrier(channel=0)
        yield lambda: check(h.ge2e_b200_step_rows_peers(p.e_hat.data_ptr(), p.c_hat_all.data_ptr(), p.cos_diag.data_ptr(),
                                                        nl, N, off, M, D, w.data_ptr(), b.data_ptr(), p.eps, p.variant,
                                                        p.precision, p.grad_out.data_ptr(), p.row_stat.data_ptr(),
                                                        p.row_kstar.data_ptr(), p.row_aux.data_ptr(),
                                                        p.row_scale.data_ptr(), p.red.data_ptr(), p.dE_hat.data_ptr(),
                                                        p._dC_owner, p._world, ws, p._ws_bytes, s), "step_rows_peers")
        yield lambda: (check(h.ge2e_b200_peer_publish(p.red.data_ptr(), p._red_rows, p._world, 0, 4, None, 0, s), "scalars"),
                       p._hd.barrier(channel=0))
        yield lambda: torch.sum(p.red_all, dim=0, out=p.red_sum)
        yield lambda: check(h.ge2e_b200_bwd_finalize(E.data_ptr(), p.dE_hat.data_ptr(), p.dC_local.data_ptr(),
                                                     p.cos_diag.data_ptr(), p.row_stat.data_ptr(), p.row_aux.data_ptr(),
                                                     p.row_scale.data_ptr(), nl, M, D, w.data_ptr(), b.data_ptr(), p.eps,
                                                     p.variant, p.grad_out.data_ptr(), p.dE.data_ptr(), s), "bwd_finalize")

    def stages():
        if p.peer:
            yield from stages_peer()
            return
        s = torch.cuda.current_stream(dev).cuda_stream
        yield lambda: check(h.ge2e_b200_prep(E.data_ptr(), nl, M, D, p.precision, p.e_hat.data_ptr(),
                                             p.c_hat_mine.data_ptr(), p.cos_diag.data_ptr(), p.red.data_ptr(), s), "prep")
        yield lambda: dist.all_gather_into_tensor(p.c_hat_all, p.c_hat_mine)
        yield lambda: check(h.ge2e_b200_step_rows(p.e_hat.data_ptr(), p.c_hat_all.data_ptr(), p.cos_diag.data_ptr(), nl, N,
                                                  off, M, D, w.data_ptr(), b.data_ptr(), p.eps, p.variant, p.precision,
                                                  p.grad_out.data_ptr(), p.row_stat.data_ptr(), p.row_kstar.data_ptr(),
                                                  p.row_aux.data_ptr(), p.row_scale.data_ptr(), p.red.data_ptr(),
                                                  p.dE_hat.data_ptr(), p.dC_partial.data_ptr(), ws, p._ws_bytes, s),
                            "step_rows")
        yield lambda: dist.reduce_scatter_tensor(p.dC_local, p.dC_partial, op=dist.ReduceOp.SUM)
        yield lambda: dist.all_reduce(p.red, op=dist.ReduceOp.SUM)
        yield lambda: check(h.ge2e_b200_bwd_finalize(E.data_ptr(), p.dE_hat.data_ptr(), p.dC_local.data_ptr(),
                                                     p.cos_diag.data_ptr(), p.row_stat.data_ptr(), p.row_aux.data_ptr(),
                                                     p.row_scale.data_ptr() if p._scaled else None, nl, M, D,
                                                     w.data_ptr(), b.data_ptr(), p.eps, p.variant, p.grad_out.data_ptr(),
                                                     p.dE.data_ptr(), s), "bwd_finalize")

    rows = []
    for it in range(a.iters + 5):
        dist.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True)]
        evs[0].record()
        for fn in stages():
            fn()
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append(e)
        torch.cuda.synchronize()
        if it >= 5:
            rows.append([evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(len(names))])
    med = torch.tensor(np.median(np.asarray(rows), axis=0), device=dev, dtype=torch.float64)
    dist.all_reduce(med, op=dist.ReduceOp.MAX)
    # the same step as one CUDA graph (what bench.py times)
    gr = p.capture(E, w, b, steps=10)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10 * 1e3)
    tg = torch.tensor([float(np.median(ts))], device=dev, dtype=torch.float64)
    dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    if rank == 0:
        out = {"shape": [N, M, D], "world": world, "peer_memory": p.peer, "peer_error": p.peer_error,
               "multicast": bool(getattr(p, "_mcast", 0)), "stage_us_eager_max_over_ranks": dict(zip(names, [round(v, 1) for v in med.tolist()])),
               "sum_of_stages_us": round(float(med.sum()), 1), "graph_step_us_max_over_ranks": round(tg.item(), 1)}
        print(json.dumps(out), flush=True)
    del gr
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
