"""CPU: host-side arithmetic of the Python mirror that runs before / after the device calls (no GPU, no library
calls): the interpolant and vanishing polynomial of batch_verify_kzg against the oracle's restatement of
Polynomial::interpolate / from_monomials (polynomial.rs:177-212), and the G2 wire format."""
import random

import myzkp_oracle as o
from myzkp_oracle import Fr

from myzkp_b200.context import g2_from_bytes, g2_to_bytes
from myzkp_b200.kzg import _from_monomials, _interpolate

R = o.R_MOD


def test_interpolate_and_from_monomials_match_the_oracle():
    rnd = random.Random(3)
    for k in (1, 2, 3, 5):
        xs = rnd.sample(range(1, 1000), k) if k > 1 else [7]
        xs = [x if rnd.random() < 0.5 else (R - x) for x in xs]
        ys = [rnd.randrange(R) for _ in range(k)]
        ip = _interpolate(xs, ys)
        exp = o.Polynomial.interpolate([Fr(x) for x in xs], [Fr(y) for y in ys]).canonical()
        assert [v % R for v in ip][: len(exp)] == exp and not any(ip[len(exp):])
        z = _from_monomials(xs)
        assert z == o.Polynomial.from_monomials([Fr(x) for x in xs]).canonical()
        for x, y in zip(xs, ys):
            assert sum(c * pow(x, i, R) for i, c in enumerate(ip)) % R == y
            assert sum(c * pow(x, i, R) for i, c in enumerate(z)) % R == 0


def test_g2_wire_format_round_trip():
    pt = o.g2_fast_mul(424242)
    assert g2_from_bytes(g2_to_bytes(pt)) == pt
    assert g2_to_bytes(pt) == o.g2_to_bytes(pt)
    assert g2_from_bytes(bytes(128)) is None and g2_to_bytes(None) == bytes(128)
