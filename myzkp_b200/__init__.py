"""myzkp_b200: B200-native KZG prover hot path (BN128) behind MyZKP's own
setup_kzg / commit_kzg / open_kzg / commit_gemini surface.

The work is done by libmyzkp_b200.so (hand-written CUDA for sm_100a, C ABI in
include/myzkp_b200.h).  There is no CPU fallback.
"""
from ._lib import MyzkpError  # noqa: F401
from .context import Context, MultiContext, R_MOD, P_MOD  # noqa: F401
from .kzg import (  # noqa: F401
    BN128, BatchProofKZG, CommitmentKZG, G1Point, G2Point, Polynomial, ProofDegreeBound, ProofKZG, PublicKeyKZG,
    batch_open_kzg, commit_kzg, open_kzg, prove_degree_bound, setup_kzg, setup_kzg_with_full_g2,
    batch_verify_kzg, optimal_ate_pairing, verify_degree_bound, verify_kzg,
)
from .gemini import ProofGemini, SplitFoldError, commit_gemini, open_gemini, split_and_fold_commit, verify_gemini  # noqa: F401
