"""GPU parity: device field arithmetic and G1 group law vs the oracle, through the C ABI."""
import random

import numpy as np
import pytest

import myzkp_oracle as o
from myzkp_oracle import _fast_add

pytestmark = pytest.mark.gpu
P, R = o.P_MOD, o.R_MOD


def _ints(arr):
    return [int.from_bytes(arr[i].tobytes(), "little") for i in range(arr.shape[0])]


@pytest.mark.parametrize("field,m", [(0, P), (1, R)])
def test_field_ops_random_and_edges(ctx, field, m):
    rnd = random.Random(11 + field)
    edge = [0, 1, 2, m - 1, m - 2, (1 << 256) % m, (1 << 255) % m, m // 2, 0xFFFFFFFF, (1 << 224) - 1, (1 << 253)]
    a = edge * len(edge) + [rnd.randrange(m) for _ in range(1 << 14)]
    b = [e for e in edge for _ in edge] + [rnd.randrange(m) for _ in range(1 << 14)]
    for op, fn in [(0, lambda x, y: (x + y) % m), (1, lambda x, y: (x - y) % m), (2, lambda x, y: x * y % m)]:
        got = _ints(ctx.test_field_op(field, op, a, b))
        assert got == [fn(x, y) for x, y in zip(a, b)], f"op {op}"
    small = a[:256]
    assert _ints(ctx.test_field_op(field, 3, small)) == [pow(x, -1, m) if x else 0 for x in small]  # inverse(0)=0
    assert _ints(ctx.test_field_op(field, 5, a[:2048])) == [pow(x, -1, m) if x else 0 for x in a[:2048]]  # binary GCD
    assert _ints(ctx.test_field_op(field, 4, small)) == [(-x) % m for x in small]


def test_field_kats_on_device(ctx):
    # cuda/test_fr.cu:19-54 style KATs: 5*7, (-2)(-12) = 24; field.rs small cases lifted to Fr
    a = [5, R - 2, R - 1, 0]
    b = [7, R - 12, R - 1, 12345]
    assert _ints(ctx.test_field_op(1, 2, a, b)) == [35, 24, 1, 0]
    assert _ints(ctx.test_field_op(1, 0, [R - 1], [1])) == [0]
    assert _ints(ctx.test_field_op(1, 1, [0], [1])) == [R - 1]


def test_g1_anchors(ctx):  # bn128.rs:285-301 on the device
    g = (1, 2)
    two_g = ctx.test_g1_op(1, [g])[0]
    assert two_g == (0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3,
                     0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4)
    k = lambda v: int(v).to_bytes(32, "little") + bytes(32)
    res = ctx.test_g1_op(2, [g] * 6, [k(R), k(0), k(1), k(9), k(5), k(R - 1)])
    assert res[0] is None and res[1] is None and res[2] == g
    assert ctx.test_g1_op(0, [res[3]], [res[4]])[0] == o.fast_mul(14)
    assert res[5] == (1, P - 2)


def test_g1_ops_vs_oracle_including_special_cases(ctx):
    rnd = random.Random(3)
    pts = [o.fast_mul(rnd.randrange(1, R)) for _ in range(24)] + [(1, 2), None]
    neg = lambda p: None if p is None else (p[0], (-p[1]) % P)
    a, b = [], []
    for x in pts:
        for y in (rnd.choice(pts), x, neg(x), None):
            a.append(x)
            b.append(y)
    exp = [_fast_add(x, y) for x, y in zip(a, b)]
    assert ctx.test_g1_op(0, a, b) == exp  # mixed add incl. P+P, P+(-P), inf
    assert ctx.test_g1_op(3, a, b) == exp  # XYZZ + XYZZ
    assert ctx.test_g1_op(1, pts) == [_fast_add(x, x) for x in pts]
    ks = [rnd.randrange(R) for _ in pts]
    got = ctx.test_g1_op(2, pts, [int(k).to_bytes(32, "little") + bytes(32) for k in ks])
    for p, k_, g_ in zip(pts, ks, got):
        assert g_ == (None if p is None else o.fast_mul(k_, p))
    # the faithful (reference-structured) oracle on a few
    gp = o.generator_g1()
    assert ctx.test_g1_op(2, [(1, 2)], [int(12345).to_bytes(32, "little") + bytes(32)])[0] == (gp * 12345).affine_ints()
