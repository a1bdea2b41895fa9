"""Stage timeline of the single-kernel small-batch step (CTA 0, SM clocks).

    python scripts/small_step_trace.py --build      # here: nvcc the instrumented variant into build/variants/
    python scripts/small_step_trace.py cfg1 cfg2    # on the GPU box: load that variant, print the timeline

The production library has no instrumentation: the variant is ge2e_simt.cu compiled with -DGE2E_DEBUG_BUILD
(the kernel then stamps clock() after every stage into the otherwise unused row_kstar buffer when the
environment says GE2E_SMALL_STOP=99) linked with the regular objects.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANT = os.path.join(ROOT, "build", "variants", "libge2e_small_trace.so")
CSRC = os.path.join(ROOT, "speaker_embedding_ge2e_loss_b200", "csrc")

if "--build" in sys.argv:
    os.makedirs(os.path.dirname(VARIANT), exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
             "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    obj = "/tmp/ge2e_simt_trace.o"
    subprocess.check_call(["nvcc", *flags, "-DGE2E_DEBUG_BUILD", "-c", os.path.join(CSRC, "ge2e_simt.cu"), "-o", obj])
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", VARIANT, obj,
                           *[os.path.join(CSRC, f) for f in ("ge2e_api.o", "ge2e_tc.o", "ge2e_tail.o")]])
    print("built", VARIANT)
    sys.exit(0)

os.environ["GE2E_SMALL_STOP"] = "99"
import torch  # noqa: E402

sys.path.insert(0, ROOT)
from speaker_embedding_ge2e_loss_b200 import _lib  # noqa: E402

_lib.LIB_PATH = VARIANT          # before the first lib() call
from bench import WORKLOADS, make_batch  # noqa: E402
from speaker_embedding_ge2e_loss_b200 import GE2EPlan  # noqa: E402

dev = torch.device("cuda:0")
_lib.lib().ge2e_b200_debug_small_step(2)      # trace every supported shape, not only the selected ones
NAMES = ["start", "prep", "barrier1", "load c_hat/e_hat", "cos block", "softmax rows", "dE_hat rows", "dC shares",
         "barrier2", "dC gather", "finalize"]
for wl in ([a for a in sys.argv[1:] if not a.startswith("-")] or ["cfg2"]):
    N, M, D = WORKLOADS[wl]
    E = make_batch(N, M, D).to(dev)
    w = torch.tensor(10.0, device=dev); b = torch.tensor(-5.0, device=dev)
    plan = GE2EPlan(N, M, D, "softmax", "fp32", device=dev)
    for _ in range(300):
        plan.step(E, w, b)
    torch.cuda.synchronize()
    st = plan.row_kstar[:11].cpu().numpy().astype("int64") & 0xffffffff
    print(wl, "loss", plan.loss.item())
    for i in range(1, 11):
        print(f"  {NAMES[i]:18s} {((st[i] - st[i - 1]) & 0xffffffff):8d} clk")
    print(f"  {'total':18s} {((st[10] - st[0]) & 0xffffffff):8d} clk")
