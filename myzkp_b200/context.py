"""Device context: owns the stream, the resident SRS table and all scratch."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def scalars_to_bytes(values) -> np.ndarray:
    """ints (or an (n,32) uint8 / (n,4) uint64 / (n,8) uint32 array) -> contiguous (n,32) uint8, LE."""
    if isinstance(values, np.ndarray):
        a = np.ascontiguousarray(values)
        if a.dtype == np.uint8 and a.ndim == 2 and a.shape[1] == 32:
            return a
        if a.dtype == np.uint64 and a.ndim == 2 and a.shape[1] == 4:
            return a.view(np.uint8).reshape(-1, 32)
        if a.dtype == np.uint32 and a.ndim == 2 and a.shape[1] == 8:
            return a.view(np.uint8).reshape(-1, 32)
        raise TypeError(f"unsupported scalar array {a.dtype} {a.shape}")
    buf = b"".join(int(v).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint8).reshape(-1, 32).copy() if buf else np.zeros((0, 32), np.uint8)


def bytes_to_int(b) -> int:
    return int.from_bytes(bytes(b), "little")


def point_from_bytes(b) -> Optional[tuple]:
    b = bytes(b)
    if b == bytes(64):
        return None
    return int.from_bytes(b[:32], "little"), int.from_bytes(b[32:], "little")


def point_to_bytes(pt) -> bytes:
    if pt is None:
        return bytes(64)
    return int(pt[0]).to_bytes(32, "little") + int(pt[1]).to_bytes(32, "little")


def g2_to_bytes(pt) -> bytes:
    """G2 wire form: x.c0 | x.c1 | y.c0 | y.c1 (32 B little-endian canonical each); infinity = 128 zero bytes."""
    if pt is None:
        return bytes(128)
    (x0, x1), (y0, y1) = pt
    return b"".join((int(v) % P_MOD).to_bytes(32, "little") for v in (x0, x1, y0, y1))


def g2_from_bytes(b):
    b = bytes(b)
    v = [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(4)]
    return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class Context:
    """One per GPU (one process per GPU in multi-GPU runs)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        code = self._lib.myzkp_ctx_create(ctypes.byref(h), int(device))
        if code != 0:
            raise _lib.MyzkpError(code, f"cannot create a CUDA context on device {device} (no CPU fallback exists)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self._lib.myzkp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, code):
        _lib.check(self.h, code)

    # -- plumbing
    def set_stream(self, cuda_stream: int):
        self._ck(self._lib.myzkp_ctx_set_stream(self.h, ctypes.c_void_p(cuda_stream)))

    def sync(self):
        """Waits for the ctx stream; raises for a non-canonical scalar seen by an asynchronous device-pointer call
        since the last synchronising call, or for a peer exchange that timed out."""
        self._ck(self._lib.myzkp_ctx_sync(self.h))

    def reserve(self, n_max: int):
        """Size all scratch for commits / opens of up to n_max coefficients now (needed before sharded calls when
        several ranks share one device)."""
        self._ck(self._lib.myzkp_ctx_reserve(self.h, int(n_max)))

    def set_table_windows(self, windows=()):
        """Restrict the next SRS table to the rows the given MSM windows need (() = automatic, 70 rows)."""
        mask = 0
        for c in windows:
            mask |= 1 << int(c)
        self._ck(self._lib.myzkp_ctx_set_table_windows(self.h, mask))

    def table_info(self):
        rows, nbytes, win = ctypes.c_int(0), ctypes.c_uint64(0), ctypes.c_uint32(0)
        self._ck(self._lib.myzkp_srs_table_info(self.h, ctypes.byref(rows), ctypes.byref(nbytes), ctypes.byref(win)))
        return {"rows": rows.value, "bytes": nbytes.value, "windows": [c for c in range(1, 25) if (win.value >> c) & 1]}

    def set_msm_params(self, window_bits: int = 0, segment_len: int = 0):
        self._ck(self._lib.myzkp_ctx_set_msm_params(self.h, window_bits, segment_len))

    def set_baa_rounds(self, rounds: int = -1):
        self._ck(self._lib.myzkp_ctx_set_baa_rounds(self.h, rounds))

    def set_upload_chunks(self, chunks: int = 0):
        self._ck(self._lib.myzkp_ctx_set_upload_chunks(self.h, chunks))

    def enable_phase_timing(self, on: bool = True):
        self._ck(self._lib.myzkp_ctx_enable_phase_timing(self.h, 1 if on else 0))

    def msm_phases(self, back: int = 0):
        """({phase: ms}, info) of the MSM `back` calls ago (0 = last); syncs on its last event."""
        ms = (ctypes.c_float * 5)()
        info = (ctypes.c_uint64 * 6)()
        self._ck(self._lib.myzkp_ctx_msm_phases(self.h, back, ms, info))
        names = ("recode", "sort", "accumulate", "merge_heads", "bucket_reduce")
        keys = ("window_bits", "windows", "entries", "segment_len", "segments", "buckets")
        return {k: float(v) for k, v in zip(names, ms)}, {k: int(v) for k, v in zip(keys, info)}

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """Pinned host buffer as a uint8 numpy array (freed with host_free)."""
        p = ctypes.c_void_p()
        code = self._lib.myzkp_host_alloc(ctypes.byref(p), nbytes)
        if code != 0:
            raise _lib.MyzkpError(code, "pinned allocation failed")
        buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p
        return arr

    def host_free(self, arr: np.ndarray):
        p = self._pinned.pop(arr.ctypes.data)
        self._lib.myzkp_host_free(p)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.myzkp_kernel_launches(self.h))

    # -- SRS
    def srs_generate(self, alpha: int, n: int, first: int = 0):
        a = np.frombuffer(int(alpha % R_MOD).to_bytes(32, "little"), dtype=np.uint8).copy()
        self._ck(self._lib.myzkp_srs_generate_g1(self.h, _ptr(a), first, n))

    def srs_generate_g2(self, alpha: int, n: int, first: int = 0, base=None):
        """[alpha^(first+i)] base for i < n on G2 as ((x.c0, x.c1), (y.c0, y.c1)) tuples / None for infinity;
        base = None is BN128::generator_g2()."""
        a = np.frombuffer(int(alpha % R_MOD).to_bytes(32, "little"), dtype=np.uint8).copy()
        b = None
        if base is not None:
            b = np.frombuffer(g2_to_bytes(base), dtype=np.uint8).copy()
        out = np.zeros((n, 128), np.uint8)
        self._ck(self._lib.myzkp_srs_generate_g2(self.h, _ptr(a), _ptr(b) if b is not None else None, first, n, _ptr(out)))
        return [g2_from_bytes(out[i]) for i in range(n)]

    def srs_load(self, points: Sequence):
        """points: list of (x, y) / None, or an (n,64) uint8 array."""
        if isinstance(points, np.ndarray):
            a = np.ascontiguousarray(points, dtype=np.uint8).reshape(-1, 64)
        else:
            a = np.frombuffer(b"".join(point_to_bytes(p) for p in points), dtype=np.uint8).reshape(-1, 64).copy()
        self._ck(self._lib.myzkp_srs_load_g1(self.h, _ptr(a), a.shape[0]))

    def srs_read(self, off: int, n: int):
        out = np.zeros((n, 64), np.uint8)
        self._ck(self._lib.myzkp_srs_read_g1(self.h, off, n, _ptr(out)))
        return [point_from_bytes(out[i]) for i in range(n)]

    @property
    def srs_len(self) -> int:
        return int(self._lib.myzkp_srs_len(self.h))

    # -- commit / open on host buffers
    def commit(self, coefs):
        a = scalars_to_bytes(coefs)
        out = np.zeros(64, np.uint8)
        self._ck(self._lib.myzkp_kzg_commit(self.h, _ptr(a), a.shape[0], _ptr(out)))
        return point_from_bytes(out)

    def commit_batch(self, polys):
        """commit_gemini (gemini.rs:112-114): all polynomials in one call; the small ones share one MSM pipeline."""
        arrs = [scalars_to_bytes(p) for p in polys]
        k = len(arrs)
        if k == 0:
            return []
        ptrs = (ctypes.c_void_p * k)(*[a.ctypes.data if a.shape[0] else None for a in arrs])
        lens = (ctypes.c_size_t * k)(*[a.shape[0] for a in arrs])
        out = np.zeros((k, 64), np.uint8)
        self._ck(self._lib.myzkp_kzg_commit_batch(self.h, ptrs, lens, k, _ptr(out)))
        return [point_from_bytes(out[i]) for i in range(k)]

    def open(self, coefs, u: int):
        a = scalars_to_bytes(coefs)
        ub = np.frombuffer(int(u).to_bytes(32, "little"), dtype=np.uint8).copy()
        y = np.zeros(32, np.uint8)
        w = np.zeros(64, np.uint8)
        self._ck(self._lib.myzkp_kzg_open(self.h, _ptr(a), a.shape[0], _ptr(ub), _ptr(y), _ptr(w)))
        return bytes_to_int(y), point_from_bytes(w)

    def batch_open(self, coefs, us):
        a = scalars_to_bytes(coefs)
        ub = scalars_to_bytes([int(u) for u in us])
        k = ub.shape[0]
        ys = np.zeros((k, 32), np.uint8)
        w = np.zeros(64, np.uint8)
        self._ck(self._lib.myzkp_kzg_batch_open(self.h, _ptr(a), a.shape[0], _ptr(ub) if k else None, k,
                                                _ptr(ys) if k else None, _ptr(w)))
        return [bytes_to_int(ys[i]) for i in range(k)], point_from_bytes(w)

    def prove_degree_bound(self, coefs, d: int):
        a = scalars_to_bytes(coefs)
        out = np.zeros(64, np.uint8)
        self._ck(self._lib.myzkp_kzg_prove_degree_bound(self.h, _ptr(a), a.shape[0], d, _ptr(out)))
        return point_from_bytes(out)

    def g1_msm(self, scalars, points=None):
        """sum_i scalars[i] * points[i]; points None = the resident SRS."""
        a = scalars_to_bytes(scalars)
        pb = None
        if points is not None:
            pb = np.frombuffer(b"".join(point_to_bytes(p) for p in points), dtype=np.uint8).reshape(-1, 64).copy()
            if pb.shape[0] != a.shape[0]:
                raise ValueError("scalars and points differ in length")
        out = np.zeros(64, np.uint8)
        self._ck(self._lib.myzkp_g1_msm(self.h, _ptr(a), _ptr(pb) if pb is not None else None, a.shape[0], _ptr(out)))
        return point_from_bytes(out)

    def g2_msm(self, scalars, points):
        """sum_i scalars[i] * points[i] over G2 (accumulate_curve_points, zksnark/utils.rs:83-93); points as
        ((x.c0, x.c1), (y.c0, y.c1)) tuples / None."""
        a = scalars_to_bytes(scalars)
        pb = np.frombuffer(b"".join(g2_to_bytes(p) for p in points), dtype=np.uint8).reshape(-1, 128).copy()
        if pb.shape[0] != a.shape[0]:
            raise ValueError("scalars and points differ in length")
        out = np.zeros(128, np.uint8)
        self._ck(self._lib.myzkp_g2_msm(self.h, _ptr(a), _ptr(pb), a.shape[0], _ptr(out)))
        return g2_from_bytes(out)

    def pairing(self, g1_points, g2_points):
        """[e(P_i, Q_i)] (optimal_ate_pairing, bn128.rs:147-181), each as the 12 Fq coefficients of w^k."""
        n = len(g1_points)
        if len(g2_points) != n:
            raise ValueError("g1 and g2 point lists differ in length")
        a = np.frombuffer(b"".join(point_to_bytes(p) for p in g1_points), dtype=np.uint8).copy()
        b = np.frombuffer(b"".join(g2_to_bytes(p) for p in g2_points), dtype=np.uint8).copy()
        out = np.zeros((n, 12, 32), np.uint8)
        self._ck(self._lib.myzkp_pairing(self.h, _ptr(a), _ptr(b), n, _ptr(out)))
        return [[bytes_to_int(out[i, k]) for k in range(12)] for i in range(n)]

    def pairing_product_is_one(self, g1_points, g2_points) -> bool:
        """prod_i e(P_i, Q_i) == 1 with one final exponentiation."""
        n = len(g1_points)
        if len(g2_points) != n:
            raise ValueError("g1 and g2 point lists differ in length")
        a = np.frombuffer(b"".join(point_to_bytes(p) for p in g1_points), dtype=np.uint8).copy()
        b = np.frombuffer(b"".join(g2_to_bytes(p) for p in g2_points), dtype=np.uint8).copy()
        res = ctypes.c_int(0)
        self._ck(self._lib.myzkp_pairing_product_is_one(self.h, _ptr(a), _ptr(b), n, ctypes.byref(res)))
        return bool(res.value)

    def fr_eval(self, coefs, u: int) -> int:
        a = scalars_to_bytes(coefs)
        ub = np.frombuffer(int(u).to_bytes(32, "little"), dtype=np.uint8).copy()
        y = np.zeros(32, np.uint8)
        self._ck(self._lib.myzkp_fr_eval(self.h, _ptr(a), a.shape[0], _ptr(ub), _ptr(y)))
        return bytes_to_int(y)

    def fr_quotient(self, coefs, u: int):
        a = scalars_to_bytes(coefs)
        n = a.shape[0]
        ub = np.frombuffer(int(u).to_bytes(32, "little"), dtype=np.uint8).copy()
        y = np.zeros(32, np.uint8)
        q = np.zeros((max(n - 1, 0), 32), np.uint8)
        self._ck(self._lib.myzkp_fr_quotient(self.h, _ptr(a), n, _ptr(ub), _ptr(y), _ptr(q)))
        return bytes_to_int(y), q

    def gemini_fold_commit(self, coefs, rhos, want_folds: bool = False):
        a = scalars_to_bytes(coefs)
        n = a.shape[0]
        r = scalars_to_bytes(rhos)
        m = r.shape[0]
        out = np.zeros((m + 1, 64), np.uint8)
        folds = np.zeros((max(n - 1, 0), 32), np.uint8) if want_folds else None
        self._ck(
            self._lib.myzkp_gemini_fold_commit(
                self.h, _ptr(a), n, _ptr(r) if m else None, m, _ptr(out), _ptr(folds) if want_folds and n > 1 else None
            )
        )
        pts = [point_from_bytes(out[i]) for i in range(m + 1)]
        return (pts, folds) if want_folds else pts

    # -- test hooks
    def test_field_op(self, field: int, op: int, a, b=None) -> np.ndarray:
        aa = scalars_to_bytes(a)
        bb = scalars_to_bytes(b) if b is not None else None
        out = np.zeros_like(aa)
        self._ck(self._lib.myzkp_test_field_op(self.h, field, op, _ptr(aa), _ptr(bb) if bb is not None else None, _ptr(out), aa.shape[0]))
        return out

    def test_g1_op(self, op: int, a, b=None):
        aa = np.frombuffer(b"".join(point_to_bytes(p) for p in a), dtype=np.uint8).reshape(-1, 64).copy()
        bb = None
        if b is not None:
            bb = np.frombuffer(b"".join(x if isinstance(x, (bytes, bytearray)) else point_to_bytes(x) for x in b), dtype=np.uint8).reshape(-1, 64).copy()
        out = np.zeros_like(aa)
        self._ck(self._lib.myzkp_test_g1_op(self.h, op, _ptr(aa), _ptr(bb) if bb is not None else None, _ptr(out), aa.shape[0]))
        return [point_from_bytes(out[i]) for i in range(aa.shape[0])]


# -- device-pointer entry points (used by bench.py and the multi-GPU layer) ----
def _dev_methods():
    def commit_dev(self, d_coefs: int, n: int, d_out64: int):
        """Asynchronous on the ctx stream; pointers are raw device addresses (e.g. tensor.data_ptr())."""
        self._ck(self._lib.myzkp_kzg_commit_dev(self.h, ctypes.c_void_p(d_coefs), n, ctypes.c_void_p(d_out64)))

    def open_dev(self, d_coefs: int, n: int, u: int, d_out_y32: int, d_out_w64: int):
        ub = np.frombuffer(int(u).to_bytes(32, "little"), dtype=np.uint8).copy()
        self._ck(self._lib.myzkp_kzg_open_dev(self.h, ctypes.c_void_p(d_coefs), n, _ptr(ub), ctypes.c_void_p(d_out_y32),
                                              ctypes.c_void_p(d_out_w64)))

    def msm_partial_dev(self, d_scalars: int, n: int, srs_off: int, d_out_xyzz128: int):
        self._ck(self._lib.myzkp_g1_msm_partial_dev(self.h, ctypes.c_void_p(d_scalars), n, srs_off,
                                                    ctypes.c_void_p(d_out_xyzz128)))

    def msm_partial_host(self, scalars: np.ndarray, srs_off: int, d_out_xyzz128: int):
        """Host (ideally pinned) scalars -> XYZZ partial on the device; asynchronous on the ctx stream."""
        a = scalars_to_bytes(scalars)
        self._keepalive = a  # the upload is asynchronous
        self._ck(self._lib.myzkp_g1_msm_partial(self.h, _ptr(a), a.shape[0], srs_off, ctypes.c_void_p(d_out_xyzz128)))

    def sum_partials_dev(self, d_partials: int, k: int, d_out64: int):
        self._ck(self._lib.myzkp_g1_sum_partials_dev(self.h, ctypes.c_void_p(d_partials), k, ctypes.c_void_p(d_out64)))

    def fr_range_eval_dev(self, d_coefs: int, n: int, u: int, d_out_h32: int, d_out_upow32: int):
        ub = np.frombuffer(int(u).to_bytes(32, "little"), dtype=np.uint8).copy()
        self._ck(self._lib.myzkp_fr_range_eval_dev(self.h, ctypes.c_void_p(d_coefs), n, _ptr(ub),
                                                   ctypes.c_void_p(d_out_h32), ctypes.c_void_p(d_out_upow32)))

    def fr_range_quotient_dev(self, d_coefs: int, n: int, u: int, carry_in: int, d_q: int, d_c0: int):
        ub = np.frombuffer(int(u).to_bytes(32, "little"), dtype=np.uint8).copy()
        cb = np.frombuffer(int(carry_in).to_bytes(32, "little"), dtype=np.uint8).copy()
        self._ck(self._lib.myzkp_fr_range_quotient_dev(self.h, ctypes.c_void_p(d_coefs), n, _ptr(ub), _ptr(cb),
                                                       ctypes.c_void_p(d_q), ctypes.c_void_p(d_c0)))

    # ---- exchange fused over peer memory (csrc/peer.cu) ----
    def peer_export(self) -> bytes:
        """Allocate / reset this rank's exchange buffer and return its 64-byte CUDA IPC handle."""
        h = np.zeros(64, dtype=np.uint8)
        self._ck(self._lib.myzkp_peer_export(self.h, _ptr(h)))
        return h.tobytes()

    def peer_attach(self, rank: int, world: int, handles: bytes):
        """handles: the world's 64-byte handles concatenated in rank order."""
        if len(handles) != 64 * world:
            raise ValueError("need one 64-byte handle per rank")
        a = np.frombuffer(handles, dtype=np.uint8).copy()
        self._ck(self._lib.myzkp_peer_attach(self.h, rank, world, _ptr(a)))

    def peer_attach_local(self, rank: int, ctxs):
        """Contexts of one process (each already peer_export()ed), in rank order."""
        arr = (ctypes.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        self._ck(self._lib.myzkp_peer_attach_local(self.h, rank, len(ctxs), arr))

    def peer_detach(self):
        self._ck(self._lib.myzkp_peer_detach(self.h))

    def peer_set_timeout_ms(self, ms: int):
        self._ck(self._lib.myzkp_peer_set_timeout_ms(self.h, ms))

    def commit_sharded_dev(self, d_scalars: int, n_local: int, d_out_c64: int):
        self._ck(self._lib.myzkp_kzg_commit_sharded_dev(self.h, ctypes.c_void_p(d_scalars), n_local, ctypes.c_void_p(d_out_c64)))

    def exchange_sum_dev(self, d_partial_xyzz128: int, d_out_c64: int):
        self._ck(self._lib.myzkp_g1_exchange_sum_dev(self.h, ctypes.c_void_p(d_partial_xyzz128), ctypes.c_void_p(d_out_c64)))

    def commit_sharded(self, scalars: np.ndarray):
        """Host scalars of this rank's range -> the whole polynomial's commitment (synchronous)."""
        a = scalars_to_bytes(scalars)
        out = np.zeros(64, dtype=np.uint8)
        self._ck(self._lib.myzkp_kzg_commit_sharded(self.h, _ptr(a), a.shape[0], _ptr(out)))
        return point_from_bytes(out.tobytes())

    def open_sharded_dev(self, d_coefs: int, n_local: int, u: int, d_out_y32: int, d_out_w64: int):
        ub = np.frombuffer(int(u).to_bytes(32, "little"), dtype=np.uint8).copy()
        self._ck(self._lib.myzkp_kzg_open_sharded_dev(self.h, ctypes.c_void_p(d_coefs), n_local, _ptr(ub),
                                                      ctypes.c_void_p(d_out_y32), ctypes.c_void_p(d_out_w64)))

    def open_sharded(self, coefs: np.ndarray, u: int):
        """Host coefficients of this rank's range -> (y, W) of the whole polynomial (synchronous)."""
        a = scalars_to_bytes(coefs)
        ub = np.frombuffer(int(u).to_bytes(32, "little"), dtype=np.uint8).copy()
        y = np.zeros(32, np.uint8)
        w = np.zeros(64, np.uint8)
        self._ck(self._lib.myzkp_kzg_open_sharded(self.h, _ptr(a), a.shape[0], _ptr(ub), _ptr(y), _ptr(w)))
        return bytes_to_int(y), point_from_bytes(w)

    for f in (commit_dev, open_dev, msm_partial_dev, msm_partial_host, sum_partials_dev, fr_range_eval_dev, fr_range_quotient_dev,
              peer_export, peer_attach, peer_attach_local, peer_detach, peer_set_timeout_ms, commit_sharded_dev,
              exchange_sum_dev, commit_sharded, open_sharded_dev, open_sharded):
        setattr(Context, f.__name__, f)


_dev_methods()


class _RankView(Context):
    """A rank of a MultiContext: same methods as Context, but the handle is owned by the multi-device context."""

    def __init__(self, lib, handle, device):
        self._lib = lib
        self.h = ctypes.c_void_p(handle)
        self.device = device

    def close(self):
        self.h = None


class MultiContext:
    """One process driving several GPUs through the C ABI alone (csrc/multi.cu): the SRS is range-sharded over the
    listed devices, commit / open hand every device its coefficient slice and finish in the peer-memory exchange
    kernel.  No torch.distributed, no NCCL.  A device may be listed more than once (ranks sharing a GPU)."""

    def __init__(self, devices: Sequence[int]):
        self._lib = _lib.load()
        ids = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        h = ctypes.c_void_p()
        code = self._lib.myzkp_mctx_create(ctypes.byref(h), ids, len(devices))
        if code != 0:
            raise _lib.MyzkpError(code, f"cannot create contexts on devices {list(devices)} (no CPU fallback exists)")
        self.h = h
        self.devices = list(devices)

    def close(self):
        if getattr(self, "h", None):
            self._lib.myzkp_mctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, code):
        if code != 0:
            msg = self._lib.myzkp_mctx_last_error(self.h)
            raise _lib.MyzkpError(code, msg.decode() if msg else "")

    @property
    def world(self) -> int:
        return int(self._lib.myzkp_mctx_world(self.h))

    @property
    def srs_len(self) -> int:
        return int(self._lib.myzkp_mctx_srs_len(self.h))

    def rank(self, g: int) -> Context:
        return _RankView(self._lib, self._lib.myzkp_mctx_rank(self.h, g), self.devices[g])

    def srs_generate(self, alpha: int, n: int):
        a = np.frombuffer(int(alpha % R_MOD).to_bytes(32, "little"), dtype=np.uint8).copy()
        self._ck(self._lib.myzkp_mctx_srs_generate_g1(self.h, _ptr(a), n))

    def srs_load(self, points):
        if isinstance(points, np.ndarray):
            a = np.ascontiguousarray(points, dtype=np.uint8).reshape(-1, 64)
        else:
            a = np.frombuffer(b"".join(point_to_bytes(p) for p in points), dtype=np.uint8).reshape(-1, 64).copy()
        self._ck(self._lib.myzkp_mctx_srs_load_g1(self.h, _ptr(a), a.shape[0]))

    def commit(self, coefs):
        a = scalars_to_bytes(coefs)
        out = np.zeros(64, np.uint8)
        self._ck(self._lib.myzkp_mctx_kzg_commit(self.h, _ptr(a), a.shape[0], _ptr(out)))
        return point_from_bytes(out)

    def open(self, coefs, u: int):
        a = scalars_to_bytes(coefs)
        ub = np.frombuffer(int(u).to_bytes(32, "little"), dtype=np.uint8).copy()
        y = np.zeros(32, np.uint8)
        w = np.zeros(64, np.uint8)
        self._ck(self._lib.myzkp_mctx_kzg_open(self.h, _ptr(a), a.shape[0], _ptr(ub), _ptr(y), _ptr(w)))
        return bytes_to_int(y), point_from_bytes(w)
