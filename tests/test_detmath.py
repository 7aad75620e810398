"""csrc/f184_detmath.h (host build, through the oracle's hook): honest sin/cos/log/log2/exp2/pow within 2 ulp of
numpy's float32 results on the ranges the path uses, and exact IEEE half conversions."""
import ctypes as C

import numpy as np


def run(oracle_lib, op, x, y=None):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), np.float32)
    out = np.empty_like(x)
    rc = oracle_lib.dll.f184o_debug_detmath(C.c_uint32(op), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_void_p(out.ctypes.data), C.c_size_t(x.size))
    assert rc == 0
    return out


def ulp_diff(a, b):
    a, b = a.astype(np.float32), b.astype(np.float32)
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia)
    ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    return np.abs(ia - ib)


def test_sin_cos_on_the_hash_range(oracle_lib):
    rng = np.random.default_rng(1)
    x = rng.uniform(-3e5, 3e5, 200000).astype(np.float32)     # fract(sin(x) * 43758.5453) feeds on |x| up to ~2e5
    for op, f in ((0, np.sin), (1, np.cos)):
        got, want = run(oracle_lib, op, x), f(x.astype(np.float64)).astype(np.float32)
        big = np.abs(want) > 1e-3
        assert ulp_diff(got[big], want[big]).max() <= 2
        assert np.abs(got - want).max() < 2e-7


def test_log_exp_pow(oracle_lib):
    rng = np.random.default_rng(2)
    x = rng.uniform(1e-6, 4096.0, 100000).astype(np.float32)
    assert ulp_diff(run(oracle_lib, 2, x), np.log(x.astype(np.float64)).astype(np.float32))[np.abs(np.log(x)) > 1e-2].max() <= 2
    assert ulp_diff(run(oracle_lib, 3, x), np.log2(x.astype(np.float64)).astype(np.float32))[np.abs(np.log2(x)) > 1e-2].max() <= 2
    e = rng.uniform(-100, 20, 100000).astype(np.float32)
    assert ulp_diff(run(oracle_lib, 4, e), np.exp2(e.astype(np.float64)).astype(np.float32)).max() <= 2
    c = (np.arange(256) / 255.0).astype(np.float32)           # the decode pow(colour, 2.2), indirect.frag:157
    got, want = run(oracle_lib, 5, c, np.full_like(c, 2.2)), np.power(c.astype(np.float64), np.float32(2.2).astype(np.float64))
    assert np.abs(got - want).max() < 4e-7 and got[0] == 0.0 and abs(got[255] - 1.0) < 2e-7


def test_half_round_trip_is_ieee(oracle_lib):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-70000, 70000, 100000), rng.uniform(-1e-4, 1e-4, 100000), [0.0, -0.0, 65504.0, 65520.0, 1e-8, np.inf, -np.inf]]).astype(np.float32)
    got = run(oracle_lib, 6, x)
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
