"""Synthetic inputs of SURVEY section 8(d): uniform scalars in [0, r) from
numpy PCG64, as (n, 4) little-endian uint64 limbs (= n x 32 B on the wire)."""
from __future__ import annotations

import numpy as np

from .context import R_MOD

_R_LIMBS = [(R_MOD >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]

SEED_SCALARS = 0xC0FFEE  # + log2 N
SEED_ALPHA = 0x5EED
SEED_OPEN = 0xBEEF
SEED_GEMINI_COEF = 0x6E1
SEED_GEMINI_RHO = 0x6E2


def _lt_r(a: np.ndarray) -> np.ndarray:
    """row-wise a < r for (n,4) uint64 limbs."""
    lt = np.zeros(a.shape[0], dtype=bool)
    eq = np.ones(a.shape[0], dtype=bool)
    for i in (3, 2, 1, 0):
        lt |= eq & (a[:, i] < np.uint64(_R_LIMBS[i]))
        eq &= a[:, i] == np.uint64(_R_LIMBS[i])
    return lt


def random_scalars(n: int, seed: int) -> np.ndarray:
    """n uniform scalars in [0, r): draw 4 x u64, mask to 254 bits, reject >= r."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n, 4), dtype=np.uint64)
    filled = 0
    while filled < n:
        m = max(1024, int((n - filled) * 1.35))
        a = rng.integers(0, 2**64, size=(m, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 62) - 1)
        a = a[_lt_r(a)]
        k = min(a.shape[0], n - filled)
        out[filled : filled + k] = a[:k]
        filled += k
    return out


def random_scalar(seed: int) -> int:
    return limbs_to_ints(random_scalars(1, seed))[0]


def limbs_to_ints(a: np.ndarray):
    a = np.ascontiguousarray(a)
    raw = a.view(np.uint8).reshape(a.shape[0], 32)
    return [int.from_bytes(raw[i].tobytes(), "little") for i in range(a.shape[0])]


def ints_to_limbs(vals) -> np.ndarray:
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def horner_mod_r(coefs_limbs: np.ndarray, x: int) -> int:
    """f(x) mod r for (n,4) uint64 coefficients - the O(N) expected-value path."""
    acc = 0
    for c in reversed(limbs_to_ints(coefs_limbs)):
        acc = (acc * x + c) % R_MOD
    return acc
