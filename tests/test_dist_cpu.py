"""CPU, world_size 2 over gloo: the range-sharding arithmetic and collective plumbing of
myzkp_b200.dist (commit and open), with the oracle standing in for the GPU group ops."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import myzkp_oracle as o
from myzkp_b200.dist import ShardedKZG, compose_carries, shard_range

R, P = o.R_MOD, o.P_MOD


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 9, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(lo <= hi for lo, hi in spans)


def test_compose_carries_matches_synthetic_division():
    import random
    rnd = random.Random(4)
    coefs = [rnd.randrange(R) for _ in range(23)]
    u = rnd.randrange(R)
    for world in (1, 2, 3, 5):
        spans = [shard_range(len(coefs), r, world) for r in range(world)]
        hs = [sum(coefs[lo + i] * pow(u, i, R) for i in range(hi - lo)) % R for lo, hi in spans]
        ms = [pow(u, hi - lo, R) for lo, hi in spans]
        carries = compose_carries(hs, ms)
        _, q = o.synthetic_division(coefs, u)
        for (lo, hi), c in zip(spans, carries):
            # carry entering range = q_{hi-1} (0 for the top range)
            assert c == (q[hi - 1] if hi - 1 < len(q) else 0)


class OracleOps:
    """Stands in for DeviceOps: local SRS = [alpha^(lo+i)]G, group ops by the oracle."""

    def __init__(self, alpha, lo):
        self.alpha, self.lo = alpha, lo

    def msm_partial(self, scalars, n):
        k = sum(s * pow(self.alpha, self.lo + i, R) for i, s in enumerate(scalars[:n])) % R
        pt = o.fast_mul(k)
        b = o.point_to_bytes(pt) + bytes(64)
        return torch.tensor(list(b), dtype=torch.uint8)

    def sum_partials(self, gathered, k, out64):
        raw = bytes(gathered.tolist())
        acc = None
        for r in range(k):
            acc = o._fast_add(acc, o.point_from_bytes(raw[128 * r : 128 * r + 64]))
        out64.copy_(torch.tensor(list(o.point_to_bytes(acc)), dtype=torch.uint8))

    def range_eval(self, coefs, n, u):
        h = sum(c * pow(u, i, R) for i, c in enumerate(coefs[:n])) % R
        return torch.tensor(list(o.fe_to_bytes(h) + o.fe_to_bytes(pow(u, n, R))), dtype=torch.uint8)

    def range_quotient(self, coefs, n, u, carry):
        q = [0] * n
        c = carry
        for i in range(n - 1, -1, -1):
            q[i] = c
            c = (coefs[i] + u * c) % R
        return q, torch.tensor(list(o.fe_to_bytes(c)), dtype=torch.uint8)


def _worker(rank, world, port, coefs, alpha, u, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = len(coefs)
        lo, hi = shard_range(n, rank, world)
        prover = ShardedKZG(OracleOps(alpha, lo), rank, world, n)
        out = torch.zeros(64, dtype=torch.uint8)
        prover.commit(coefs[lo:hi], out)
        y = torch.zeros(32, dtype=torch.uint8)
        w = torch.zeros(64, dtype=torch.uint8)
        prover.open(coefs[lo:hi], u, y, w)
        ret[rank] = (bytes(out.tolist()), bytes(y.tolist()), bytes(w.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 2, 1])
def test_sharded_commit_and_open_world2_gloo(n):
    import random
    rnd = random.Random(n)
    coefs = [rnd.randrange(R) for _ in range(n)]
    alpha, u = rnd.randrange(R), rnd.randrange(R)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, coefs, alpha, u, ret), nprocs=2, join=True)
    exp_c = o.point_to_bytes(o.expected_commit(coefs, alpha))
    ey, ew = o.expected_open(coefs, u, alpha)
    for rank in (0, 1):
        c, y, w = ret[rank]
        assert c == exp_c
        assert int.from_bytes(y, "little") == ey
        assert o.point_from_bytes(w) == ew
