"""Range-sharded multi-GPU commit / open: one process per GPU, torch.distributed for the exchange.

Only the MSM is sharded (BASELINE north star).  Rank g owns the contiguous index
range [lo_g, hi_g) of the SRS (its resident table holds exactly those points) and
receives only that slice of the coefficients.  Per commit each rank produces ONE
partial G1 element (128-byte XYZZ); the partials are all-gathered (NCCL over
NVLink; world_size x 128 B, latency-bound) and every rank sums them and
normalises to affine, so all ranks hold the identical commitment.

open() shards the quotient scan the same way: each rank reduces its coefficient
range to (h_g, u^{n_g}); after one all-gather of 64 B per rank every rank composes
the carry entering its range from above on the host (world_size field ops) and
runs the local scan; the quotient slice then feeds the same sharded MSM.

Exchange: by default (`DeviceOps.attach_peers`) the partials never go through a
collective library.  Each rank's 8 KiB exchange buffer is mapped into every peer
(CUDA IPC; the 64-byte handles are swapped once with torch.distributed) and the
commit / open end in ONE kernel that stores the 128-byte partial into all peers'
HBM over NVLink, waits for theirs and sums + normalises in the same launch
(csrc/peer.cu).  The all_gather(NCCL) + sum path stays as the portable route and is
what the gloo CPU tests drive.

The group-element arithmetic is behind an `ops` object: `DeviceOps` drives the
CUDA library; the CPU tests (gloo, world_size 2) plug the oracle in to check the
sharding arithmetic and the collective plumbing without a GPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

from .context import R_MOD


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ceil-split: rank g owns [g*ceil(n/G), (g+1)*ceil(n/G)) clipped to n."""
    per = -(-n // world) if n else 0
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def compose_carries(hs: Sequence[int], ms: Sequence[int]) -> List[int]:
    """carry entering each range from above.  Range g maps a carry c to h_g + m_g * c
    (m_g = u^{n_g}); carry_{G-1} = 0, carry_g = h_{g+1} + m_{g+1} * carry_{g+1}."""
    world = len(hs)
    carries = [0] * world
    c = 0
    for g in range(world - 1, -1, -1):
        carries[g] = c
        c = (hs[g] + ms[g] * c) % R_MOD
    return carries


class DeviceOps:
    """Group/field work of one rank on its GPU through the C ABI (device pointers)."""

    def __init__(self, ctx, device: torch.device):
        self.ctx = ctx
        self.device = device
        # everything below mixes library calls with torch collectives and .cpu() reads on torch's current stream:
        # put the library on that stream so the two are ordered (the library's default is a private stream)
        ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.partial = torch.zeros(128, dtype=torch.uint8, device=device)
        self.pair = torch.zeros(64, dtype=torch.uint8, device=device)  # (h, u^n)
        self.c0 = torch.zeros(32, dtype=torch.uint8, device=device)
        self.q = None

    def attach_peers(self, rank: int, world: int, group=None) -> bool:
        """Map every rank's exchange buffer (handles swapped through torch.distributed).  Returns
        False (and leaves the NCCL route in place) if any rank could not map its peers."""
        handle = self.ctx.peer_export()
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(self.device)
        allh = torch.zeros(64 * world, dtype=torch.uint8, device=self.device)
        if world > 1:
            dist.all_gather_into_tensor(allh, mine, group=group)
        else:
            allh.copy_(mine)
        ok = 1
        try:
            self.ctx.peer_attach(rank, world, allh.cpu().numpy().tobytes())
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.fused = bool(int(flag.item()))
        if not self.fused:
            self.ctx.peer_detach()
        return self.fused

    fused = False

    def commit_fused(self, d_scalars: int, n: int, out64: torch.Tensor) -> None:
        self.ctx.commit_sharded_dev(d_scalars, n, out64.data_ptr())

    def exchange_sum(self, partial: torch.Tensor, out64: torch.Tensor) -> None:
        self.ctx.exchange_sum_dev(partial.data_ptr(), out64.data_ptr())

    def open_fused(self, d_coefs: int, n: int, u: int, out_y32: torch.Tensor, out_w64: torch.Tensor) -> None:
        self.ctx.open_sharded_dev(d_coefs, n, u, out_y32.data_ptr(), out_w64.data_ptr())

    def open_single(self, d_coefs: int, n: int, u: int, out_y32: torch.Tensor, out_w64: torch.Tensor) -> None:
        """world_size 1: the plain single-GPU open (one scan, no exchange, no host round trip)."""
        self.ctx.open_dev(d_coefs, n, u, out_y32.data_ptr(), out_w64.data_ptr())

    def msm_partial(self, d_scalars: int, n: int) -> torch.Tensor:
        self.ctx.msm_partial_dev(d_scalars, n, 0, self.partial.data_ptr())
        return self.partial

    def msm_partial_host(self, host_scalars, n: int) -> torch.Tensor:
        self.ctx.msm_partial_host(host_scalars[:n], 0, self.partial.data_ptr())
        return self.partial

    def sum_partials(self, gathered: torch.Tensor, k: int, out64: torch.Tensor) -> None:
        self.ctx.sum_partials_dev(gathered.data_ptr(), k, out64.data_ptr())

    def range_eval(self, d_coefs: int, n: int, u: int) -> torch.Tensor:
        self.ctx.fr_range_eval_dev(d_coefs, n, u, self.pair.data_ptr(), self.pair.data_ptr() + 32)
        return self.pair

    def range_quotient(self, d_coefs: int, n: int, u: int, carry: int) -> Tuple[int, torch.Tensor]:
        if self.q is None or self.q.numel() < max(n, 1) * 32:
            self.q = torch.zeros(max(n, 1) * 32, dtype=torch.uint8, device=self.device)
        self.ctx.fr_range_quotient_dev(d_coefs, n, u, carry, self.q.data_ptr(), self.c0.data_ptr())
        return self.q.data_ptr(), self.c0


class ShardedKZG:
    def __init__(self, ops, rank: int, world: int, n_total: int, group=None):
        self.ops, self.rank, self.world, self.n_total, self.group = ops, rank, world, n_total, group
        self.lo, self.hi = shard_range(n_total, rank, world)
        self.n_local = self.hi - self.lo
        self._gather128 = None
        self._gather64 = None
        self._gather32 = None

    def _all_gather(self, t: torch.Tensor, cache_name: str) -> torch.Tensor:
        buf = getattr(self, cache_name)
        if buf is None or buf.device != t.device:
            buf = torch.zeros(self.world * t.numel(), dtype=torch.uint8, device=t.device)
            setattr(self, cache_name, buf)
        if self.world == 1:
            buf.copy_(t)
        else:
            dist.all_gather_into_tensor(buf, t, group=self.group)
        return buf

    def commit(self, d_scalars_local: int, out64: torch.Tensor, n_local: int = None) -> None:
        """d_scalars_local: this rank's coefficient slice [lo, hi) (device address)."""
        n = self.n_local if n_local is None else n_local
        if getattr(self.ops, "fused", False):
            return self.ops.commit_fused(d_scalars_local, n, out64)
        partial = self.ops.msm_partial(d_scalars_local, n)
        gathered = self._all_gather(partial, "_gather128")
        self.ops.sum_partials(gathered, self.world, out64)

    def commit_host(self, host_scalars_local, out64: torch.Tensor) -> None:
        """Same with this rank's coefficient slice in host (pinned) memory: upload pipelined with the MSM."""
        partial = self.ops.msm_partial_host(host_scalars_local, self.n_local)
        if getattr(self.ops, "fused", False):
            return self.ops.exchange_sum(partial, out64)
        gathered = self._all_gather(partial, "_gather128")
        self.ops.sum_partials(gathered, self.world, out64)

    def open(self, d_coefs_local: int, u: int, out_y32: torch.Tensor, out_w64: torch.Tensor) -> None:
        """open_kzg over the sharded polynomial: every rank ends with (y, W)."""
        if getattr(self.ops, "fused", False):
            return self.ops.open_fused(d_coefs_local, self.n_local, u, out_y32, out_w64)
        if self.world == 1 and hasattr(self.ops, "open_single"):
            return self.ops.open_single(d_coefs_local, self.n_local, u, out_y32, out_w64)
        pair = self.ops.range_eval(d_coefs_local, self.n_local, u)
        g = self._all_gather(pair, "_gather64").cpu().numpy().tobytes()
        hs = [int.from_bytes(g[64 * r : 64 * r + 32], "little") for r in range(self.world)]
        ms = [int.from_bytes(g[64 * r + 32 : 64 * r + 64], "little") for r in range(self.world)]
        carry = compose_carries(hs, ms)[self.rank]
        d_q, c0 = self.ops.range_quotient(d_coefs_local, self.n_local, u, carry)
        # q_{lo+i} pairs with local SRS point i; the global top coefficient q_{n-1} is 0
        partial = self.ops.msm_partial(d_q, self.n_local)
        gathered = self._all_gather(partial, "_gather128")
        self.ops.sum_partials(gathered, self.world, out_w64)
        c0s = self._all_gather(c0, "_gather32")
        out_y32.copy_(c0s[:32])  # rank 0's c_0 is y
