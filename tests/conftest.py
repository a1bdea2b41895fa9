import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "ge2e_reference_vectors.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class GoldenCase:
    def __init__(self, z, name):
        self.name = name
        self.E = z[f"{name}/E"]
        self.w, self.b = (float(v) for v in z[f"{name}/wb"])
        for tag in ("r32", "r64", "rc"):
            loss, dw, db = (float(v) for v in z[f"{name}/{tag}/scalars"])
            setattr(self, tag, dict(loss=loss, dw=dw, db=db, per=z[f"{name}/{tag}/per"],
                                    dE=z[f"{name}/{tag}/dE"]))


def golden_names():
    with np.load(GOLDEN) as z:
        return sorted({k.split("/")[0] for k in z.files})


@pytest.fixture(scope="session")
def golden():
    with np.load(GOLDEN) as z:
        return {n: GoldenCase(z, n) for n in {k.split("/")[0] for k in z.files}}
