"""ctypes binding of libvoroffset_b200.so (the C ABI in include/voroffset_b200.h).

There is no fallback: if the CUDA library has not been built, or no device is usable, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VO_LIB") or os.path.join(_PKG, "libvoroffset_b200.so")   # (VO_LIB: A/B builds, scripts/ only)

VO_OK = 0
ERRORS = {1: "VO_ERR_ARG", 2: "VO_ERR_CUDA", 3: "VO_ERR_NOMEM", 4: "VO_ERR_OVERFLOW"}

OPS3D = {"dilation": 0, "erosion": 1, "opening": 2, "closing": 3}
METHODS = {"ours": 0, "brute_force": 1}
OPS2D = {"dilate": 0, "erode": 1, "open": 2, "close": 3, "negate": 4}

_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# every symbol include/voroffset_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "vo_create", "vo_destroy", "vo_last_error", "vo_version", "vo_span_bytes", "vo_free", "vo_stream",
    "vo_launch_count", "vo_morph3d", "vo_morph3d_rows", "vo_morph2d", "vo_xor3d", "vo_dvol_upload", "vo_dvol_download",
    "vo_dvol_info", "vo_dvol_free", "vo_dvol_rows", "vo_dvol_concat_rows", "vo_morph3d_dev", "vo_xor3d_dev",
    "vo_pass1_dev", "vo_pass2_dev", "vo_dmid_free", "vo_dmid_info", "vo_morph2d_dev",
    "vo_mark", "vo_elapsed_ms", "vo_last_profile", "vo_dvol_from_device", "vo_set_option", "vo_dvol_rows_to",
    "vo_slab_begin", "vo_slab_finish", "vo_slab_abort", "vo_dexelize_dev",
    "vo_mg_create", "vo_mg_unique_id", "vo_mg_create_rank", "vo_mg_destroy", "vo_mg_world", "vo_mg_local_count",
    "vo_mg_rank", "vo_mg_ctx", "vo_mg_last_error", "vo_mg_morph3d", "vo_mg_morph3d_dev", "vo_mg_stats",
]

_lib = None


class VoroffsetError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.vo_create.argtypes = [C.c_int, C.POINTER(_vp)]
    L.vo_destroy.argtypes = [_vp]
    L.vo_destroy.restype = None
    L.vo_last_error.argtypes = [_vp]
    L.vo_last_error.restype = C.c_char_p
    L.vo_version.restype = C.c_char_p
    L.vo_free.argtypes = [_vp]
    L.vo_free.restype = None
    L.vo_stream.argtypes = [_vp]
    L.vo_stream.restype = _vp
    L.vo_launch_count.argtypes = [_vp]
    L.vo_launch_count.restype = C.c_uint64
    L.vo_morph3d.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _vp, _vp,
                             C.c_double, C.POINTER(_u32p), C.POINTER(_f64p), C.POINTER(C.c_uint64), _f64p, _f64p]
    L.vo_morph3d_rows.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _vp, _vp,
                                  C.c_double, C.c_int, C.c_int, C.POINTER(_u32p), C.POINTER(_f64p), C.POINTER(C.c_uint64),
                                  _f64p, _f64p]
    L.vo_morph2d.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_double,
                             C.POINTER(_u32p), C.POINTER(_f64p), C.POINTER(C.c_uint64), _f64p]
    L.vo_xor3d.argtypes = [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _vp, _vp, _vp, _vp,
                           C.POINTER(_u32p), C.POINTER(_f64p), C.POINTER(C.c_uint64), _f64p]
    L.vo_dvol_upload.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, C.POINTER(_vp)]
    L.vo_dvol_download.argtypes = [_vp, _vp, _vp, _vp]
    L.vo_dvol_info.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64),
                               C.POINTER(_vp), C.POINTER(_vp)]
    L.vo_dvol_free.argtypes = [_vp, _vp]
    L.vo_dvol_free.restype = None
    L.vo_dvol_rows.argtypes = [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)]
    L.vo_dvol_concat_rows.argtypes = [_vp, _vp, _vp, _vp, C.POINTER(_vp)]
    L.vo_dvol_rows_to.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.vo_morph3d_dev.argtypes = [_vp, C.c_int, C.c_int, _vp, C.c_double, C.c_double, C.c_double,
                                 C.POINTER(_vp), _f64p, _f64p]
    L.vo_xor3d_dev.argtypes = [_vp, _vp, _vp, C.c_double, C.c_double, C.c_double, C.POINTER(_vp), _f64p]
    L.vo_pass1_dev.argtypes = [_vp, _vp, C.c_double, C.POINTER(_vp), _f64p]
    L.vo_pass2_dev.argtypes = [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp), _f64p]
    L.vo_dmid_free.argtypes = [_vp, _vp]
    L.vo_dmid_free.restype = None
    L.vo_dmid_info.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    L.vo_morph2d_dev.argtypes = [_vp, C.c_int, _vp, C.c_int, C.c_double, C.POINTER(_vp), _f64p]
    L.vo_dexelize_dev.argtypes = [_vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_double, C.c_double, C.c_double,
                                  C.c_int, C.c_int, C.POINTER(_vp), _f64p]
    L.vo_set_option.argtypes = [_vp, C.c_char_p, C.c_char_p]
    L.vo_mark.argtypes = [_vp, C.c_int]
    L.vo_elapsed_ms.argtypes = [_vp, C.c_int, C.c_int, _f64p]
    L.vo_last_profile.argtypes = [_vp, _f64p, _f64p]
    L.vo_dvol_from_device.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, C.c_uint64, C.POINTER(_vp)]
    L.vo_slab_begin.argtypes = [_vp, _vp, C.c_double, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                _vp, _vp, C.c_uint64, _vp, _vp, C.c_uint64, _vp, C.POINTER(_vp)]
    L.vo_slab_finish.argtypes = [_vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp, C.c_uint64, C.POINTER(_vp), _f64p, _f64p]
    L.vo_slab_abort.argtypes = [_vp, _vp]
    L.vo_slab_abort.restype = None
    L.vo_mg_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]
    L.vo_mg_unique_id.argtypes = [_vp]
    L.vo_mg_create_rank.argtypes = [C.c_int, C.c_int, C.c_int, _vp, C.POINTER(_vp)]
    L.vo_mg_destroy.argtypes = [_vp]
    L.vo_mg_destroy.restype = None
    L.vo_mg_world.argtypes = [_vp]
    L.vo_mg_local_count.argtypes = [_vp]
    L.vo_mg_rank.argtypes = [_vp, C.c_int]
    L.vo_mg_ctx.argtypes = [_vp, C.c_int]
    L.vo_mg_ctx.restype = _vp
    L.vo_mg_last_error.argtypes = [_vp]
    L.vo_mg_last_error.restype = C.c_char_p
    L.vo_mg_morph3d.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _vp, _vp,
                                C.c_double, C.POINTER(_u32p), C.POINTER(_f64p), C.POINTER(C.c_uint64), _f64p, _f64p]
    L.vo_mg_morph3d_dev.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(_vp), C.c_double, C.c_double, C.c_double,
                                    C.POINTER(_vp), _f64p, _f64p]
    L.vo_mg_stats.argtypes = [_vp, C.c_int, _f64p, _f64p, C.POINTER(C.c_uint64), C.POINTER(C.c_int),
                              C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = L
    return L


class Context:
    """One vo_ctx: one device, one stream. Not thread-safe (include/voroffset_b200.h)."""

    def __init__(self, device: int = 0, _borrowed=None):
        self.lib = load()
        if _borrowed is not None:               # a context owned by a multi-GPU group (vo_mg_ctx): never destroyed here
            self.handle, self.device, self._owned = _vp(_borrowed), int(device), False
            return
        self._owned = True
        h = _vp()
        rc = self.lib.vo_create(int(device), C.byref(h))
        if rc != VO_OK:
            raise VoroffsetError(rc, f"vo_create(device={device}) failed: no usable CUDA device (no CPU fallback)")
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            if getattr(self, "_owned", True):
                self.lib.vo_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != VO_OK:
            raise VoroffsetError(rc, self.lib.vo_last_error(self.handle).decode(errors="replace"))

    @property
    def launches(self) -> int:
        return int(self.lib.vo_launch_count(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.vo_stream(self.handle) or 0)

    def set_option(self, key: str, value: str):
        self.check(self.lib.vo_set_option(self.handle, key.encode(), value.encode()))

    def mark(self, slot: int):
        self.check(self.lib.vo_mark(self.handle, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double(0)
        self.check(self.lib.vo_elapsed_ms(self.handle, a, b, C.byref(ms)))
        return ms.value

    def last_profile(self):
        a, b = C.c_double(0), C.c_double(0)
        self.check(self.lib.vo_last_profile(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def take_host(self, poff, pspans, nlists: int, nspans: int):
        """Copy a (vo_free-able) result pair into numpy arrays and release the pinned blocks."""
        off = np.ctypeslib.as_array(poff, shape=(nlists + 1,)).copy()
        if nspans:
            spans = np.ctypeslib.as_array(pspans, shape=(2 * nspans,)).copy().reshape(-1, 2)
        else:
            spans = np.zeros((0, 2))
        self.lib.vo_free(C.cast(poff, _vp))
        self.lib.vo_free(C.cast(pspans, _vp))
        return off, spans


def ptr(a: np.ndarray) -> int:
    return a.ctypes.data


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
