"""The C-ABI library on a machine WITHOUT a GPU: it loads, exports every symbol the header declares, answers host-only
queries, and refuses to create handles (there is no CPU path).  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "fbe_cabi.h")).read()
    return sorted(set(re.findall(r"FBE_API\s+[\w\s\*]+?\b(fbe_\w+)\s*\(", txt)))


def test_header_symbols_exported(fbe):
    syms = declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(fbe, s), f"{s} declared in include/fbe_cabi.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", fbe._name], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fbe_\w+)", out))
    assert exported == set(syms), (exported ^ set(syms))


def test_no_torch_or_oracle_dependency(fbe):
    out = subprocess.run(["ldd", fbe._name], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "opencv" not in out


def test_plan_query_matches_survey_table(fbe):
    from fishbirdeyevisualslam_b200._lib import ExtractorCfg, ptr
    cfg = ExtractorCfg(1000, 1.2, 8, 15, 5, 1, 0)
    out = np.zeros((8, 8), np.int32)
    assert fbe.fbe_plan_query(C.byref(cfg), 480, 640, ptr(out), None, None, None, None) == 0
    assert out[:, 0].tolist() == [640, 533, 444, 370, 309, 257, 214, 179]
    assert out[:, 1].tolist() == [480, 400, 333, 278, 231, 193, 161, 134]
    assert out[:, 6].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert (out[:, 2] * out[:, 3]).sum() == 815 and out[0, 7] == 1
    cfg = ExtractorCfg(8000, 1.2, 12, 15, 5, 1, 0)
    out = np.zeros((12, 8), np.int32)
    assert fbe.fbe_plan_query(C.byref(cfg), 2160, 3840, ptr(out), None, None, None, None) == 0
    assert out[:, 6].tolist() == [1502, 1251, 1043, 869, 724, 604, 503, 419, 349, 291, 243, 202]
    assert (out[:, 2] * out[:, 3]).sum() == 27939


def test_plan_query_matches_oracle_tables(fbe, oracle):
    from fishbirdeyevisualslam_b200._lib import ExtractorCfg, ptr
    for nf, sf, nl, (h, w) in [(2000, 1.2, 8, (720, 1280)), (1000, 1.2, 8, (384, 384)), (500, 1.5, 3, (300, 500)), (777, 1.1, 10, (400, 950))]:
        cfg = ExtractorCfg(nf, sf, nl, 15, 5, 1, 0)
        out = np.zeros((nl, 8), np.int32)
        sc, isc, s2, is2 = (np.zeros(nl, np.float32) for _ in range(4))
        assert fbe.fbe_plan_query(C.byref(cfg), h, w, ptr(out), ptr(sc), ptr(isc), ptr(s2), ptr(is2)) == 0
        o = oracle.OracleExtractor(nf, sf, nl, 15, 5)
        t = o.tables()
        assert np.array_equal(sc, t["scale"]) and np.array_equal(isc, t["inv_scale"])
        assert np.array_equal(s2, t["sigma2"]) and np.array_equal(is2, t["inv_sigma2"])
        assert np.array_equal(out[:, 6], t["per_level"])
        o(np.zeros((h, w), np.uint8))
        assert [tuple(r) for r in out[:, :2].tolist()] == [o.level_size(l) for l in range(nl)]


def test_unsupported_geometry_is_refused(fbe):
    from fishbirdeyevisualslam_b200._lib import ExtractorCfg, ptr, FBE_E_UNSUPPORTED
    out = np.zeros((8, 8), np.int32)
    cfg = ExtractorCfg(1000, 1.2, 8, 15, 5, 1, 0)
    assert fbe.fbe_plan_query(C.byref(cfg), 100, 100, ptr(out), None, None, None, None) == FBE_E_UNSUPPORTED   # top level < one FAST cell
    assert fbe.fbe_plan_query(C.byref(cfg), 5000, 5000, ptr(out), None, None, None, None) == FBE_E_UNSUPPORTED
    assert fbe.fbe_plan_query(C.byref(ExtractorCfg(1000, 1.2, 2, 15, 5, 1, 0)), 900, 200, ptr(out), None, None, None, None) == FBE_E_UNSUPPORTED  # 0 roots


def test_hamming_host_inline(fbe, oracle):
    rng = np.random.default_rng(0)
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    for _ in range(200):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        d = ORBmatcher.DescriptorDistance(a, b)
        assert d == oracle.hamming256(a, b) == int(np.unpackbits(a ^ b).sum())


def test_no_cpu_fallback_without_gpu(fbe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fishbirdeyevisualslam_b200 import _lib
    from fishbirdeyevisualslam_b200.extractor import ORBextractor
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    with pytest.raises(_lib.FbeError) as e:
        ORBextractor(1000, 1.2, 8, 15, 5)
    assert e.value.code == _lib.FBE_E_CUDA
    with pytest.raises(_lib.FbeError):
        ORBmatcher(0.9, True)


def test_product_never_imports_oracle():
    """The product path must not import, include, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "fishbirdeyevisualslam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cc", ".cpp", ".h", ".hpp")):
                continue
            for ln in open(os.path.join(dp, f), errors="ignore").read().splitlines():
                if "oracle" in ln.lower():
                    assert not re.search(r"\bimport\b|#include|CDLL|dlopen", ln), f"{f}: {ln.strip()}"
