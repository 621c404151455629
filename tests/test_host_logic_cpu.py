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


def test_cpp_mirror_scalar_negation():
    """include/myzkp_b200.hpp negates scalars on their wire bytes for the verifier equations: r - x, 0 -> 0."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "hpp_host_check")
    src = os.path.join(root, "tests", "cpp", "hpp_host_check.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(root, "include"), "-o", exe, src,
                           "-L", os.path.join(root, "myzkp_b200"), "-lmyzkp_b200",
                           "-Wl,-rpath," + os.path.join(root, "myzkp_b200")])
    rnd = random.Random(9)
    xs = [0, 1, 2, R - 1, R - 2, 1 << 200, 0xFF, (R + 1) // 2] + [rnd.randrange(R) for _ in range(20)]
    out = subprocess.run([exe] + ["%064x" % x for x in xs], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v, 16) for v in out] == [(-x) % R for x in xs]
