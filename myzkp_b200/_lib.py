"""ctypes loader for libmyzkp_b200.so (the C ABI in include/myzkp_b200.h).

There is no CPU fallback: if the CUDA library is missing this raises, and if
no GPU is usable ctx creation raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MYZKP_B200_LIB: development knob to load an experimental build of the same library (still CUDA-only)
LIB_PATH = os.environ.get("MYZKP_B200_LIB") or os.path.join(_HERE, "libmyzkp_b200.so")

MYZKP_OK = 0
ERR_NAMES = {-1: "INVALID_ARG", -2: "NONCANONICAL", -3: "CUDA", -4: "OOM", -5: "NO_SRS"}

# every symbol include/myzkp_b200.h declares: name -> (restype, argtypes)
_c = ctypes
_vp, _sz, _i, _u8p = _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_char_p
SIGNATURES = {
    "myzkp_ctx_create": (_i, [_c.POINTER(_vp), _i]),
    "myzkp_ctx_destroy": (_i, [_vp]),
    "myzkp_ctx_set_stream": (_i, [_vp, _vp]),
    "myzkp_ctx_sync": (_i, [_vp]),
    "myzkp_ctx_reserve": (_i, [_vp, _sz]),
    "myzkp_last_error": (_c.c_char_p, [_vp]),
    "myzkp_kernel_launches": (_c.c_uint64, [_vp]),
    "myzkp_ctx_set_msm_params": (_i, [_vp, _i, _i]),
    "myzkp_ctx_set_baa_rounds": (_i, [_vp, _i]),
    "myzkp_ctx_set_upload_chunks": (_i, [_vp, _i]),
    "myzkp_ctx_enable_phase_timing": (_i, [_vp, _i]),
    "myzkp_ctx_msm_phases": (_i, [_vp, _i, _c.POINTER(_c.c_float), _c.POINTER(_c.c_uint64)]),
    "myzkp_host_alloc": (_i, [_c.POINTER(_vp), _sz]),
    "myzkp_host_free": (_i, [_vp]),
    "myzkp_srs_load_g1": (_i, [_vp, _vp, _sz]),
    "myzkp_srs_generate_g1": (_i, [_vp, _vp, _sz, _sz]),
    "myzkp_srs_read_g1": (_i, [_vp, _sz, _sz, _vp]),
    "myzkp_srs_generate_g2": (_i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "myzkp_g2_msm": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "myzkp_pairing": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "myzkp_pairing_product_is_one": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "myzkp_srs_len": (_sz, [_vp]),
    "myzkp_ctx_set_table_windows": (_i, [_vp, _c.c_uint32]),
    "myzkp_srs_table_info": (_i, [_vp, _c.POINTER(_i), _c.POINTER(_c.c_uint64), _c.POINTER(_c.c_uint32)]),
    "myzkp_kzg_commit": (_i, [_vp, _vp, _sz, _vp]),
    "myzkp_kzg_open": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "myzkp_kzg_commit_batch": (_i, [_vp, _c.POINTER(_vp), _c.POINTER(_sz), _sz, _vp]),
    "myzkp_gemini_fold_commit": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "myzkp_kzg_batch_open": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "myzkp_kzg_prove_degree_bound": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "myzkp_g1_msm": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "myzkp_fr_eval": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "myzkp_fr_quotient": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "myzkp_kzg_commit_dev": (_i, [_vp, _vp, _sz, _vp]),
    "myzkp_kzg_open_dev": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "myzkp_g1_msm_partial_dev": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "myzkp_g1_msm_partial": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "myzkp_peer_export": (_i, [_vp, _vp]),
    "myzkp_peer_attach": (_i, [_vp, _i, _i, _vp]),
    "myzkp_peer_attach_local": (_i, [_vp, _i, _i, _vp]),
    "myzkp_peer_detach": (_i, [_vp]),
    "myzkp_peer_set_timeout_ms": (_i, [_vp, ctypes.c_uint32]),
    "myzkp_kzg_commit_sharded_dev": (_i, [_vp, _vp, _sz, _vp]),
    "myzkp_g1_exchange_sum_dev": (_i, [_vp, _vp, _vp]),
    "myzkp_kzg_commit_sharded": (_i, [_vp, _vp, _sz, _vp]),
    "myzkp_kzg_open_sharded_dev": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "myzkp_kzg_open_sharded": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "myzkp_device_count": (_i, []),
    "myzkp_mctx_create": (_i, [_c.POINTER(_vp), _c.POINTER(_i), _i]),
    "myzkp_mctx_destroy": (_i, [_vp]),
    "myzkp_mctx_last_error": (_c.c_char_p, [_vp]),
    "myzkp_mctx_world": (_i, [_vp]),
    "myzkp_mctx_rank": (_vp, [_vp, _i]),
    "myzkp_mctx_srs_len": (_sz, [_vp]),
    "myzkp_mctx_srs_generate_g1": (_i, [_vp, _vp, _sz]),
    "myzkp_mctx_srs_load_g1": (_i, [_vp, _vp, _sz]),
    "myzkp_mctx_kzg_commit": (_i, [_vp, _vp, _sz, _vp]),
    "myzkp_mctx_kzg_open": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "myzkp_g1_sum_partials_dev": (_i, [_vp, _vp, _sz, _vp]),
    "myzkp_fr_range_eval_dev": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "myzkp_fr_range_quotient_dev": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "myzkp_test_field_op": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz]),
    "myzkp_test_g1_op": (_i, [_vp, _i, _vp, _vp, _vp, _sz]),
    "myzkp_test_set_sort_group_cap": (_i, [_i]),
}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C myzkp_b200/csrc`). myzkp_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MyzkpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"myzkp_b200 error {code} ({ERR_NAMES.get(code, '?')}): {msg}")
        self.code = code


def check(ctx_handle, code: int) -> None:
    if code != MYZKP_OK:
        msg = load().myzkp_last_error(ctx_handle)
        raise MyzkpError(code, msg.decode() if msg else "")
