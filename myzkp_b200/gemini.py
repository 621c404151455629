"""Host-side mirror of gemini.rs:51-114 (split_and_fold + commit_gemini)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .context import R_MOD, bytes_to_int
from .kzg import (BatchProofKZG, G1Point, Polynomial, PublicKeyKZG, batch_open_kzg, batch_verify_kzg, prove_degree_bound,
                  verify_degree_bound)


class SplitFoldError(ValueError):
    """gemini.rs:15-32."""


def _check(n: int, n_rhos: int):
    if n == 0 or (n & (n - 1)):
        raise SplitFoldError(f"coefs.len() must be a power of two, but got {n}")
    log2_n = n.bit_length() - 1
    if n_rhos != log2_n:
        raise SplitFoldError(f"points.len() must be {log2_n}, but got {n_rhos}")


def split_and_fold_commit(coef, rhos: Sequence[int], pk: PublicKeyKZG, want_folds: bool = False):
    """split_and_fold (gemini.rs:51-103) fused with commit_gemini (gemini.rs:112-114):
    the folds stay in HBM and each level is committed where it lies."""
    n = len(coef)
    _check(n, len(rhos))
    wire = coef if isinstance(coef, np.ndarray) else [int(c) % R_MOD for c in coef]
    res = pk.ctx.gemini_fold_commit(wire, [int(r) % R_MOD for r in rhos], want_folds=want_folds)
    if want_folds:
        pts, folds = res
        polys, off, ln = [], 0, n // 2
        while ln >= 1:
            polys.append(Polynomial([bytes_to_int(folds[off + i]) for i in range(ln)]))
            off += ln
            ln //= 2
        return [G1Point._from_tuple(p) for p in pts], polys
    return [G1Point._from_tuple(p) for p in res]


def commit_gemini(polys: Sequence[Polynomial], pk: PublicKeyKZG) -> List[G1Point]:
    """gemini.rs:112-114: a commit_kzg per polynomial - here one batched call (myzkp_kzg_commit_batch)."""
    return [G1Point._from_tuple(t) for t in pk.ctx.commit_batch([p._wire() for p in polys])]


class ProofGemini:
    """gemini.rs:107-110."""

    def __init__(self, es: List[BatchProofKZG], degree_proofs: List[G1Point]):
        self.es = es
        self.degree_proofs = degree_proofs


def open_gemini(polys: Sequence[Polynomial], beta: int, pk: PublicKeyKZG) -> ProofGemini:
    """gemini.rs:116-144: every polynomial but the last is opened at (beta, -beta, beta^2);
    polynomial i gets a degree-bound proof for 2^(num_polys - i - 1)."""
    num_polys = len(polys)
    beta = int(beta) % R_MOD
    us = [beta, (-beta) % R_MOD, beta * beta % R_MOD]
    es = [batch_open_kzg(p, us, pk) for p in polys[: num_polys - 1]]
    degree_proofs = [prove_degree_bound(p, pk, 2 ** (num_polys - i - 1)) for i, p in enumerate(polys)]
    return ProofGemini(es, degree_proofs)


def verify_gemini(rhos: Sequence[int], mu: int, beta: int, commitment: Sequence[G1Point], proof: ProofGemini,
                  pk: PublicKeyKZG) -> bool:
    """gemini.rs:146-203: degree bounds of every fold, the three-point openings, then the folding identity
    2 beta e_hat_j == beta (e_j + e_neg_j) + rho_j (e_j - e_neg_j)."""
    log2_n = len(rhos)
    if log2_n != len(commitment) - 1:
        return False
    if not all(verify_degree_bound(c, p, pk, 2 ** (log2_n - i)) for i, (c, p) in enumerate(zip(commitment, proof.degree_proofs))):
        return False
    beta = int(beta) % R_MOD
    us = [beta, (-beta) % R_MOD, beta * beta % R_MOD]
    if not all(batch_verify_kzg(us, c, p, pk) for c, p in zip(commitment[:-1], proof.es)):
        return False
    es = [p.ys[0] for p in proof.es]
    es_neg = [p.ys[1] for p in proof.es]
    es_hat = [p.ys[2] for p in proof.es][1:] + [int(mu) % R_MOD]
    return all((2 * beta * es_hat[j]) % R_MOD == (beta * (es[j] + es_neg[j]) + int(rhos[j]) * (es[j] - es_neg[j])) % R_MOD
               for j in range(log2_n))
