"""ctypes binding of libisle_cuda.so (include/isle_cuda.h).

There is no CPU path: importing this module without the built library, or creating a
context without a B200-class GPU, raises.  Build with ``python -m isle_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libisle_cuda.so")

ISLE_OK, ISLE_ERR_CUDA, ISLE_ERR_ARG, ISLE_ERR_RANGE, ISLE_ERR_NOCONV, ISLE_ERR_NOGPU = range(6)


class IsleCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libisle_cuda error {code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m isle_b200.build` "
            "(isle_b200 has no CPU or PyTorch fallback)")
    # libisle_cuda needs soname libnccl.so.2.  Preload the NCCL that ships with torch (if any) so
    # that this process ends up with ONE NCCL whether torch is imported before or after us.
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
    except Exception:
        pass
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _u64, _i64, _int, _f32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_float
_pp = C.POINTER(C.c_void_p)

# name -> argtypes (restype is int everywhere except destroy / last_error)
SIGNATURES = {
    "isle_cuda_create": [_pp, _int],
    "isle_cuda_create_sharded": [_pp, _int, _int, _int, _vp],
    "isle_cuda_nccl_unique_id": [_vp],
    "isle_cuda_create_multi": [_pp, _int, _vp],
    "isle_cuda_upload_A": [_vp, _u64, _u64, _i64, _vp, _vp, _vp, _f32, _u64],
    "isle_cuda_upload_A_u32": [_vp, _u64, _u64, _i64, _vp, _vp, _vp, _f32, _u64],
    "isle_cuda_ingest_text": [_vp, _vp, _u64, _u64, _u64, _i64, C.POINTER(_i64), C.POINTER(_f32), C.POINTER(_u64), C.POINTER(_u64)],
    "isle_cuda_upload_counts": [_vp, _u64, _u64, _i64, _vp, _vp, _vp, C.POINTER(_f32), C.POINTER(_u64)],
    "isle_cuda_download_A": [_vp, _vp, _vp, _vp],
    "isle_cuda_thresholds": [_vp, _u64, _vp, C.POINTER(_i64)],
    "isle_cuda_build_B": [_vp, _vp, C.POINTER(_i64), C.POINTER(_u64)],
    "isle_cuda_sampling_weights": [_vp, _vp],
    "isle_cuda_download_B": [_vp, _vp, _vp, _vp, _vp],
    "isle_cuda_download_B_begin": [_vp, _vp, _vp, _vp, _vp],
    "isle_cuda_download_B_end": [_vp],
    "isle_cuda_frobenius": [_vp, C.POINTER(_f32)],
    "isle_cuda_spsptr_multiply": [_vp, _int, _vp, _vp],
    "isle_cuda_block_ks": [_vp, _u64, _int, _int, _f32, _u64, _vp, _vp, C.POINTER(_int)],
    "isle_cuda_set_U": [_vp, _u64, _vp],
    "isle_cuda_project": [_vp, _vp, _vp],
    "isle_cuda_kmeanspp": [_vp, _u64, _u64, _vp, _vp, C.POINTER(_f32)],
    "isle_cuda_lloyd_projected": [_vp, _u64, _vp, _int, _vp, C.POINTER(C.c_double), C.POINTER(_int)],
    "isle_cuda_assign_projected": [_vp, _u64, _vp, _vp],
    "isle_cuda_update_min_dist": [_vp, _u64, _vp, _vp],
    "isle_cuda_lift_centers": [_vp, _u64, _vp, _u64, _vp],
    "isle_cuda_sample_docs": [_vp, C.c_float, _u64, _vp, C.POINTER(C.c_uint64)],
    "isle_cuda_rth_highest_element": [_vp, _u64, _vp, _u64, _vp],
    "isle_cuda_catchword_thresholds": [_vp, _u64, _u64, _vp, _vp],
    "isle_cuda_find_catchwords": [_vp, _u64, _vp, C.c_double, _vp],
    "isle_cuda_construct_topic_model": [_vp, _u64, _vp, _vp, _u64, _vp, C.POINTER(C.c_uint64)],
    "isle_cuda_doc_topic_sums": [_vp, _vp, _vp, _vp],
    "isle_cuda_panel_products": [_vp, C.c_int64, _int, _int, _vp, _vp, _vp, _int],
    "isle_cuda_lloyd_full": [_vp, _u64, _vp, _int, _vp, C.POINTER(C.c_double), C.POINTER(_int)],
    "isle_cuda_selftest_collectives": [_vp, C.POINTER(C.c_uint64), C.POINTER(_int)],
    "isle_cuda_cleanup_eigensolver": [_vp],
    "isle_cuda_set_profiling": [_vp, _int],
    "isle_cuda_get_stat": [_vp, C.c_char_p, C.POINTER(C.c_double)],
    "isle_cuda_reset_stats": [_vp],
    "isle_cuda_set_option": [_vp, C.c_char_p, _int],
    "isle_cuda_timer_start": [_vp],
    "isle_cuda_timer_stop": [_vp, C.POINTER(C.c_double)],
}
for _n, _a in SIGNATURES.items():
    getattr(lib, _n).argtypes = _a
    getattr(lib, _n).restype = _int
lib.isle_cuda_destroy.argtypes = [_vp]
lib.isle_cuda_destroy.restype = None
lib.isle_cuda_last_error.argtypes = [_vp]
lib.isle_cuda_last_error.restype = C.c_char_p
EXPORTS = sorted(list(SIGNATURES) + ["isle_cuda_destroy", "isle_cuda_last_error"])


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """Owns one isle_cuda_ctx (one GPU, one caller thread)."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_id: bytes | None = None, n_gpus: int = 0):
        h = C.c_void_p()
        if n_gpus > 1:       # one process, n_gpus devices, one host thread per device inside the library
            rc = lib.isle_cuda_create_multi(C.byref(h), n_gpus, None)
        elif world > 1:
            buf = C.create_string_buffer(nccl_id, 128)
            rc = lib.isle_cuda_create_sharded(C.byref(h), device, rank, world, buf)
        else:
            rc = lib.isle_cuda_create(C.byref(h), device)
        if rc != ISLE_OK:
            raise IsleCudaError(rc, (lib.isle_cuda_last_error(None) or b"").decode())
        self.h = h
        self.rank, self.world = rank, world

    def call(self, name: str, *args) -> None:
        rc = getattr(lib, name)(self.h, *args)
        if rc != ISLE_OK:
            raise IsleCudaError(rc, (lib.isle_cuda_last_error(self.h) or b"").decode())

    def stat(self, name: str) -> float:
        v = C.c_double()
        self.call("isle_cuda_get_stat", name.encode(), C.byref(v))
        return v.value

    def set_option(self, name: str, value: int) -> None:
        self.call("isle_cuda_set_option", name.encode(), int(value))

    def close(self) -> None:
        if getattr(self, "h", None):
            lib.isle_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib.isle_cuda_nccl_unique_id(buf)
    if rc != ISLE_OK:
        raise IsleCudaError(rc, "ncclGetUniqueId failed")
    return buf.raw
