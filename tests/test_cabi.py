"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports exactly
what include/ge2e_b200.h declares; host-side argument checking; module surface."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ge2e_b200.h")


@pytest.fixture(scope="module")
def pkg():
    from speaker_embedding_ge2e_loss_b200 import build
    build.build()   # nvcc cross-compiles sm_100a without a GPU
    import speaker_embedding_ge2e_loss_b200 as p
    return p


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ge2e_b200_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    names = declared_functions()
    for must in ("ge2e_b200_forward", "ge2e_b200_backward", "ge2e_b200_prep", "ge2e_b200_fwd_rows",
                 "ge2e_b200_bwd_rows", "ge2e_b200_bwd_finalize", "ge2e_b200_centroids",
                 "ge2e_b200_utterance_centroids", "ge2e_b200_calc_loss", "ge2e_b200_workspace_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol(pkg):
    from speaker_embedding_ge2e_loss_b200.build import LIB_PATH
    h = ctypes.CDLL(LIB_PATH)
    for name in declared_functions():
        assert hasattr(h, name), f"{name} declared in include/ge2e_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ge2e_b200_\w+)", out))
    assert exported == set(declared_functions())


def test_binding_prototypes_cover_header(pkg):
    from speaker_embedding_ge2e_loss_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_functions()


def test_binding_prototypes_have_the_header_arity(pkg):
    """Every ctypes prototype takes exactly as many arguments as the C declaration (a missing or extra
    argument in the binding would silently shift every pointer after it)."""
    from speaker_embedding_ge2e_loss_b200 import _lib
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decls = dict(re.findall(r"\b(ge2e_b200_\w+)\s*\(([^)]*)\)", src))
    assert set(decls) == set(_lib.PROTOTYPES)
    for name, params in decls.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert len(_lib.PROTOTYPES[name][1]) == n, f"{name}: header has {n} parameters, binding {len(_lib.PROTOTYPES[name][1])}"


def test_host_side_status_codes(pkg):
    h = pkg.lib()
    assert h.ge2e_b200_version() >= 100
    assert h.ge2e_b200_strerror(0) == b"ok"
    for code in range(-6, 0):
        assert h.ge2e_b200_strerror(code) not in (b"ok", b"unknown ge2e status")
    # argument checking happens before any CUDA call: safe without a GPU
    assert h.ge2e_b200_forward(None, 4, 8, 256, None, None, 1e-6, 0, 0, None, None, None, None, None, None, None,
                               None, None, None, 0, None) == -3
    assert h.ge2e_b200_prep(1, 4, 1, 256, 0, 1, 1, 1, 1, None) == -1          # M < 2
    assert h.ge2e_b200_prep(1, 4, 8, 256, 7, 1, 1, 1, 1, None) == -3          # unknown precision
    assert h.ge2e_b200_fwd_rows(1, 1, 1, 4, 2, 0, 8, 256, 1, 1, 1e-6, 0, 0, 1, None, 1, 1, None, None, None, None,
                                None, 0, None) == -1                           # shard outside [0, n_total)
    assert h.ge2e_b200_step_rows(1, 1, 1, 4, 4, 0, 8, 256, 1, 1, 1e-6, 0, 1, 1, 1, None, 1, None, 1, 1, 1, None, 0,
                                 None) == -3                                   # row_scale is required
    assert h.ge2e_b200_calc_loss(1, 4, 8, 1e-6, 9, 1, None, None) == -3
    assert h.ge2e_b200_scale_bias_sgd(None, 1, 1, 1, 1.0, 0.01, None, None) == -3   # null parameter pointer
    assert h.ge2e_b200_scale_bias_sgd(1, 1, 1, 1, 0.0, 0.01, None, None) == -3      # max_norm must be > 0
    assert h.ge2e_b200_path(64, 64, 10, 256, 0, 0) == 0                        # fp32 -> SIMT kernels
    assert h.ge2e_b200_path(64, 64, 10, 256, 5, 0) < 0


def test_module_surface_matches_reference(pkg):
    # reference: s3_loss_function_GE2E.py:6-127 -- ctor(hp), w/b parameters, static helpers
    class General:
        device = torch.device("cpu")
        small_err = 1e-6

    class HP:
        general = General()

    crit = pkg.GE2ELoss(HP())
    names = dict(crit.named_parameters())
    assert set(names) == {"w", "b"}
    assert crit.w.item() == 10.0 and crit.b.item() == -5.0 and crit.w.dim() == 0
    assert crit.w.requires_grad and crit.b.requires_grad
    assert crit.eps == 1e-6 and crit.device == torch.device("cpu") and crit.hp is not None
    assert set(crit.state_dict()) == {"w", "b"}
    for helper in ("get_centroids", "get_cos_sim", "get_centroid", "get_utterance_centroids", "calc_loss"):
        assert callable(getattr(pkg.GE2ELoss, helper))
    # dict-style hp (utils/dict_to_dot.py builds a dict subclass)
    crit2 = pkg.GE2ELoss({"general": {"device": torch.device("cpu"), "small_err": 1e-5}})
    assert crit2.eps == 1e-5


def test_no_cpu_fallback(pkg):
    crit = pkg.GE2ELoss(None, device=torch.device("cpu"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(torch.randn(4, 8, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.GE2ELoss.get_centroids(torch.randn(4, 8, 16))
    with pytest.raises(RuntimeError):
        pkg.GE2EPlan(4, 8, 16, device="cpu")


def test_missing_library_fails_loudly(pkg, monkeypatch):
    from speaker_embedding_ge2e_loss_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libge2e_b200.so")
    with pytest.raises(_lib.GE2ELibraryError, match="no CPU"):
        _lib.lib()


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "speaker_embedding_ge2e_loss_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert "oracle" not in src, f"{fn} references oracle/"


# ------------------------------------------------------------------ step-kernel work schedule (host logic)
@pytest.mark.parametrize("u_local,n_total", [(10240, 1024), (131072, 8192), (16384, 8192), (4096, 2048), (2100, 300),
                                             (256, 8192), (300 * 7, 300), (65536, 256), (1024, 1024)])
@pytest.mark.parametrize("cg,max_cl", [(2, 74), (1, 148), (2, 8), (1, 3)])
def test_step_schedule_covers_every_pair_once(pkg, u_local, n_total, cg, max_cl):
    """Every (owner group, stream unit) pair of both passes belongs to exactly one cluster, the ranges are
    contiguous and monotone, owner groups are whole unless a cut is reported, and the busiest cluster of a
    pass carries no more than the cheapest of the three cuts allows (whole groups / one group per cluster /
    flat ranges + the cost of a straddle); pass 2 may shift up to a straddle's worth of units away from the
    clusters that finish pass 1 last."""
    import ctypes as C
    h = pkg.lib()
    de = (C.c_int * (max_cl + 1))()
    dc = (C.c_int * (max_cl + 1))()
    part = (C.c_int * 2)()
    units = (C.c_int * 4)()
    nc = h.ge2e_b200_debug_step_schedule(u_local, n_total, cg, max_cl, de, dc, part, units)
    assert 1 <= nc <= max_cl
    oge, ste, ogc, stc = list(units)
    de, dc = list(de)[:nc + 1], list(dc)[:nc + 1]
    assert de[0] == 0 and dc[0] == 0 and de[-1] == oge * ste and dc[-1] == ogc * stc
    assert all(b >= a for a, b in zip(de, de[1:])) and all(b >= a for a, b in zip(dc, dc[1:]))
    if not part[0]:
        assert all(x % ste == 0 for x in de), "whole dE_hat groups expected"
    if not part[1]:
        assert all(x % stc == 0 for x in dc), "whole dC_hat groups expected"
    straddle = 3
    for begin, og, st, slack in ((de, oge, ste, 1), (dc, ogc, stc, straddle + 2)):
        total = og * st
        loads = [begin[c + 1] - begin[c] for c in range(nc)]
        assert sum(loads) == total
        level = -(-total // nc)
        cands = [-(-og // nc) * st, level + straddle]
        if og <= nc:
            cands.append(-(-st // (nc // og)))
        assert max(loads) <= min(cands) + slack, (max(loads), cands)
        # a range never spans more than two groups unless groups are shorter than the level share
        if st >= level:
            assert all((begin[c + 1] - 1) // st - begin[c] // st <= 1 for c in range(nc) if loads[c] > 0)


def test_precision_names_resolve_without_a_gpu():
    """The Python layer's precision names -> C-ABI codes.  "fp32" asks ge2e_b200_path() whether the split-fp16
    tensor-core kernels cover the shape; without a CUDA driver (this suite's CPU runs) nothing is covered and the
    name falls back to the SIMT fp32 code -- never an exception, never a silent TF32."""
    import torch
    from speaker_embedding_ge2e_loss_b200 import _lib
    h = _lib.lib()
    got = _lib.resolve_precision("fp32", 1024, 1024, 10, 256, _lib.SOFTMAX)
    if torch.cuda.is_available():
        assert got == _lib.FP32_SPLIT and h.ge2e_b200_path(1024, 1024, 10, 256, 0, _lib.FP32_SPLIT) == 2
    else:
        assert got == _lib.FP32 and h.ge2e_b200_path(1024, 1024, 10, 256, 0, _lib.FP32_SPLIT) < 0
    assert _lib.resolve_precision("fp32", 64, 64, 10, 256, _lib.SOFTMAX) == _lib.FP32          # reference-sized batch
    assert _lib.resolve_precision("fp32", 1024, 1024, 10, 256, _lib.CONTRAST) == _lib.FP32    # contrast: SIMT backward
    assert _lib.resolve_precision("fp32_simt", 1024, 1024, 10, 256, _lib.SOFTMAX) == _lib.FP32
    assert _lib.resolve_precision("tf32", 8192, 8192, 16, 256, _lib.SOFTMAX) == _lib.TF32      # never fp16 operands unasked
    assert _lib.resolve_precision("tf32_mma", 1024, 1024, 10, 256, _lib.SOFTMAX) == _lib.TF32
    with __import__("pytest").raises(ValueError):
        _lib.resolve_precision("bf16", 64, 64, 10, 256, _lib.SOFTMAX)
    # argument checking of the new precision codes at the C level (no launch: a bad enum returns first)
    assert h.ge2e_b200_path(64, 64, 10, 256, 0, 7) < 0
    assert h.ge2e_b200_path(64, 64, 10, 256, 0, _lib.F16) < 0                 # below the tensor-core sizes
