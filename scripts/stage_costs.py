"""In-situ price of every kernel of the fwd+bwd step: CUDA-graph replay time of the full step minus
the same graph with one kernel left out (GE2E_SKIP, debug only -- results are garbage, timing is not).
    python scripts/stage_costs.py [cfg3]         (spawns one process per mask: the mask is read once)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MASKS = [("full", 0), ("-prep", 1), ("-fwd_rows", 2), ("-bwd_rows", 12), ("-finalize", 16),
         ("only prep+fwd", 28), ("only bwd kernels", 3), ("nothing (memsets)", 31)]

if len(sys.argv) > 2 and sys.argv[1] == "--child":
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from bench import L2_FLUSH_BYTES, WORKLOADS, make_batch, timed_steps
    from speaker_embedding_ge2e_loss_b200 import GE2EPlan
    N, M, D = WORKLOADS[sys.argv[2]]
    dev = torch.device("cuda:0")
    E = make_batch(N, M, D).to(dev)
    w = torch.tensor(10.0, device=dev)
    b = torch.tensor(-5.0, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    plan = GE2EPlan(N, M, D, "softmax", "tf32", device=dev)
    g = plan.capture(E, w, b)
    ms = timed_steps(g.replay, 50, 10, flush)
    print(json.dumps({"us": float(np.median(ms)) * 1e3, "us_mean": float(np.mean(ms)) * 1e3}))
    sys.exit(0)

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
base = None
for name, mask in MASKS:
    env = dict(os.environ, GE2E_SKIP=str(mask))
    out = subprocess.run([sys.executable, __file__, "--child", wl], env=env, capture_output=True, text=True)
    try:
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        print(name, "FAILED", out.stderr[-400:])
        continue
    if base is None:
        base = r["us"]
    print(f"{name:22s} step {r['us']:7.1f} us (mean {r['us_mean']:7.1f})   delta vs full {base - r['us']:6.1f} us")
