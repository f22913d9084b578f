"""CPU-only checks of the C-ABI boundary: the library builds, loads, exports every declared symbol, and its host helpers
agree with the oracle.  No kernels are launched."""
import os
import re

import numpy as np
import pytest

from hawq_b200 import _lib
from hawq_b200.build import build_library
from oracle import int_ref as ir
from tests.util import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_library()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    header = open(os.path.join(ROOT, "include", "hawq_b200.h")).read()
    declared = set(re.findall(r"\b(hawq_[a-z0-9_]+)\s*\(", header))
    declared -= {"hawq_status"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.hawq_abi_version() == 1


def test_dyadic_matches_reference_kat(lib):
    g = load_golden("kat_requant.npz")
    for r, m, e in zip(g["frexp_ratio"], g["frexp_m"], g["frexp_e"]):
        if 1 <= e <= 62:
            assert _lib.dyadic(float(r)) == (int(m), int(e))
    with pytest.raises(_lib.HawqError):
        _lib.dyadic(2.0 ** 31)
    with pytest.raises(_lib.HawqError):
        _lib.dyadic(-1.0)
    assert _lib.dyadic(1e-30) == (0, 1)


def test_host_rhe_requant_matches_oracle(lib):
    rs = np.random.RandomState(3)
    for _ in range(2000):
        v = int(rs.randint(-2 ** 31, 2 ** 31 - 1))
        m = int(rs.randint(2 ** 30, 2 ** 31 + 1))
        e = int(rs.randint(1, 63))
        want = int(np.clip(ir.rhe_shift(np.int64(v) * np.int64(m), e), -2 ** 31, 2 ** 31 - 1)) if abs(v) * m < 2 ** 62 else None
        if want is not None:
            assert lib.hawq_rhe_requant_host(v, m, e) == want
    # SURVEY A.7 ties-to-even
    m, e = _lib.dyadic(1.0 / 16)
    assert [lib.hawq_rhe_requant_host(v, m, e) for v in (8, 24, 40, -8, -24, 9, 23)] == [0, 2, 2, 0, -2, 1, 1]


def test_permute_roundtrip(lib):
    import torch
    from hawq_b200 import ops
    from tests.abi_model import unpermute_i4_weights
    w = torch.arange(2 * 3 * 64, dtype=torch.int32).remainder(127).to(torch.int8).view(2, 1, 3, 64).contiguous()
    orig = w.clone()
    ops.permute_weights_for_i4(w)
    assert not torch.equal(w, orig)
    assert np.array_equal(unpermute_i4_weights(w.numpy()), orig.numpy())


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.HawqLibraryError, match="no CPU or PyTorch fallback"):
        _lib.load()
