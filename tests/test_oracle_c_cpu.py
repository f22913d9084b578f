"""The plain-C oracle (oracle/int_ref.c) against the reference-generated KATs and the numpy oracle (oracle/int_ref.py):
two independent restatements of the same definitions must agree everywhere."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import int_ref as ir  # noqa: E402
from tests.util import load_golden  # noqa: E402

I64P, I32P, F32P = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_float)


def p64(a):
    return a.ctypes.data_as(I64P)


def p32(a):
    return a.ctypes.data_as(I32P)


@pytest.fixture(scope="module")
def clib():
    path = ge._build_oracle_c()
    assert path and os.path.isfile(path)
    lib = C.CDLL(path)
    lib.hawq_ref_dyadic.argtypes = [C.c_double, I64P, I32P]
    lib.hawq_ref_rhe_shift.restype = C.c_int64
    lib.hawq_ref_rhe_shift.argtypes = [C.c_int64, C.c_int32]
    lib.hawq_ref_requant_case0.argtypes = [I64P, C.c_int64, C.c_int32, I64P, I64P, I32P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, I64P]
    lib.hawq_ref_requant_case1.argtypes = [I64P, I64P, C.c_int64, C.c_int32, I64P, I32P, I64P, I32P, C.c_int32, C.c_int32, I64P]
    lib.hawq_ref_conv2d_nhwc.argtypes = [I32P, I32P, I64P] + [C.c_int32] * 9 + [I64P]
    lib.hawq_ref_maxpool_3x3_s2_p1.argtypes = [I64P] + [C.c_int32] * 4 + [I64P]
    lib.hawq_ref_avgpool_trunc.argtypes = [I64P] + [C.c_int32] * 3 + [I64P]
    lib.hawq_ref_quantize_input.argtypes = [F32P] + [C.c_int32] * 4 + [C.c_float, C.c_int64, C.c_int64, I64P]
    assert lib.hawq_ref_sat32_selftest() == 1
    return lib


def cdyadic(lib, r):
    m, e = C.c_int64(), C.c_int32()
    lib.hawq_ref_dyadic(float(r), C.byref(m), C.byref(e))
    return m.value, e.value


def test_batch_frexp_kat(clib):
    g = load_golden("kat_requant.npz")
    for r, m, e in zip(g["frexp_ratio"], g["frexp_m"], g["frexp_e"]):
        assert cdyadic(clib, r) == (int(m), int(e))
    rs = np.random.RandomState(0)
    for r in np.exp(rs.uniform(np.log(1e-9), np.log(1e4), 500)):
        assert cdyadic(clib, r) == ir.dyadic(r)


def test_case0_and_case1_kats_generated_by_the_reference(clib):
    g = load_golden("kat_requant.npz")
    specs = json.loads(str(g["c0_specs"]))
    for i, spec in enumerate(specs):
        acc = np.ascontiguousarray(np.moveaxis(g["c0_%d_acc" % i], 1, -1)).astype(np.int64)       # NCHW -> NHWC
        want = np.moveaxis(g["c0_%d_q" % i], 1, -1)
        ratio = ir.requant_ratio(g["c0_%d_a_sf" % i], g["c0_%d_w_sf" % i], g["c0_%d_z_sf" % i])
        m, e = zip(*[cdyadic(clib, r) for r in np.atleast_1d(ratio)])
        m, e = np.array(m, dtype=np.int64), np.array(e, dtype=np.int32)
        lo, hi = ir.clamp_range(spec["bits"], spec["mode"])
        out = np.zeros_like(acc)
        c = acc.shape[-1]
        if m.size == 1:
            m, e = np.repeat(m, c), np.repeat(e, c)
        clib.hawq_ref_requant_case0(p64(acc), acc.size // c, c, None, p64(m), p32(e), 1, 0, lo, hi, p64(out))
        assert np.array_equal(out, want), i
    for i in range(3):
        acc = np.ascontiguousarray(np.moveaxis(g["c1_%d_acc" % i], 1, -1)).astype(np.int64)
        idt = np.ascontiguousarray(np.moveaxis(g["c1_%d_id" % i], 1, -1)).astype(np.int64)
        want = np.moveaxis(g["c1_%d_q" % i], 1, -1)
        c = acc.shape[-1]
        r2 = ir.requant_ratio(g["c1_%d_a_sf" % i], g["c1_%d_w_sf" % i], g["c1_%d_z_sf" % i])
        r1 = ir.requant_ratio(g["c1_%d_id_sf" % i], g["c1_%d_id_w_sf" % i], g["c1_%d_z_sf" % i])
        m2, e2 = zip(*[cdyadic(clib, r) for r in np.broadcast_to(np.atleast_1d(r2), (c,))])
        m1, e1 = zip(*[cdyadic(clib, r) for r in np.atleast_1d(r1)])
        out = np.zeros_like(acc)
        clib.hawq_ref_requant_case1(p64(acc), p64(idt), acc.size // c, c, p64(np.array(m2, dtype=np.int64)), p32(np.array(e2, dtype=np.int32)),
                                    p64(np.array(m1, dtype=np.int64)), p32(np.array(e1, dtype=np.int32)), int(len(m1) > 1), 0, p64(out))
        assert np.array_equal(out, want), i


def test_rhe_conv_pools_and_input_against_numpy_oracle(clib):
    rs = np.random.RandomState(1)
    for _ in range(2000):
        p = int(rs.randint(-2 ** 62, 2 ** 62))
        e = int(rs.randint(1, 63))
        assert clib.hawq_ref_rhe_shift(p, e) == int(ir.rhe_shift(np.int64(p), e))
    for (n, h, w, cin, cout, k, s, pad) in [(2, 9, 9, 8, 5, 3, 1, 1), (1, 12, 10, 3, 4, 7, 2, 3), (2, 8, 8, 16, 8, 1, 2, 0), (1, 5, 7, 4, 6, 3, 2, 1)]:
        x = rs.randint(-128, 128, size=(n, h, w, cin)).astype(np.int32)
        wt = rs.randint(-128, 128, size=(cout, k, k, cin)).astype(np.int32)
        bias = rs.randint(-10 ** 6, 10 ** 6, size=cout).astype(np.int64)
        ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
        out = np.zeros((n, ho, wo, cout), dtype=np.int64)
        clib.hawq_ref_conv2d_nhwc(p32(x), p32(wt), p64(bias), n, h, w, cin, cout, k, k, s, pad, p64(out))
        assert np.array_equal(out, ir.conv2d_nhwc(x, wt, s, pad) + bias)
    x = rs.randint(-30000, 30000, size=(2, 11, 9, 6)).astype(np.int64)
    out = np.zeros((2, 6, 5, 6), dtype=np.int64)
    clib.hawq_ref_maxpool_3x3_s2_p1(p64(x), 2, 11, 9, 6, p64(out))
    assert np.array_equal(out, ir.maxpool_3x3_s2_p1(x))
    x = rs.randint(-200, 200, size=(3, 7, 7, 10)).astype(np.int64)
    x[0, :, :, 0] = -2                                   # sum = -98 = exact multiple of 49: the "+0.01" case
    x[0, :, :, 1] = 2
    out = np.zeros((3, 10), dtype=np.int64)
    clib.hawq_ref_avgpool_trunc(p64(x.reshape(3, 49, 10).copy()), 3, 49, 10, p64(out))
    assert np.array_equal(out, ir.avgpool_trunc(x, 7).reshape(3, 10))
    assert out[0, 0] == -1 and out[0, 1] == 2
    xf = rs.randn(2, 3, 6, 5).astype(np.float32)
    q = np.zeros((2, 6, 5, 3), dtype=np.int64)
    clib.hawq_ref_quantize_input(xf.ctypes.data_as(F32P), 2, 3, 6, 5, C.c_float(0.0207), -128, 127, p64(q))
    assert np.array_equal(q, ir.quantize_input(xf, np.float32(0.0207), 8, "symmetric"))
