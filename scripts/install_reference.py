"""Place the UNMODIFIED reference sources under baseline/_ref/ (git-ignored, NOT gpurun-ignored, so the
copy travels to the GPU box with the snapshot like the built .so).

    python scripts/install_reference.py [--src /root/reference]

`pip install /root/reference` does not apply: the reference has no setup.py / pyproject.toml (Pipfile
only), it is a directory of plain Python files that import each other by package path
(`embedding_model_GE2E.*`, `utils.*`).  Installing it therefore means copying those files byte for byte;
`static/` (audio, checkpoints, figures: 46 MB) is left out.  Only `tests/` and `bench.py --impl reference`
read baseline/_ref, as the thing compared WITH (the reference's own GE2ELoss / trainer / EER code running
unmodified); nothing under speaker_embedding_ge2e_loss_b200/ imports it.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
KEEP = ("embedding_model_GE2E", "utils", "strings")
TOP = ("train_embedding_model.py", "test_embedding_model.py", "README.md", "Pipfile")


def install(src: str = "/root/reference", dest: str = DEST) -> dict:
    if not os.path.isdir(src):
        raise SystemExit(f"{src} not found (the reference tree exists in the build container only)")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(dest)
    manifest = {}
    for d in KEEP:
        for base, _, files in os.walk(os.path.join(src, d)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                s = os.path.join(base, f)
                rel = os.path.relpath(s, src)
                t = os.path.join(dest, rel)
                os.makedirs(os.path.dirname(t), exist_ok=True)
                shutil.copyfile(s, t)
                manifest[rel] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    for f in TOP:
        s = os.path.join(src, f)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(dest, f))
            manifest[f] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "sha256": manifest}, fh, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    m = install(a.src)
    print(f"installed {len(m)} reference files into {DEST}")
    sys.exit(0)
