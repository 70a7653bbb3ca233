"""ctypes front end of the CPU model of the device codec (tests only)."""
import ctypes as C

import numpy as np

from . import build as _build

_L = C.CDLL(str(_build.build()))
_L.hm_inflate.restype = C.c_int
_L.hm_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
_L.hm_sub_bytes.restype = C.c_uint32
SUB = int(_L.hm_sub_bytes())


_L.hm_encode_stream_v2.restype = C.c_uint64
_L.hm_encode_stream_v2.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64)]


def encode_stream_v2(a: np.ndarray, sequential: bool = False, skip: bool = True):
    """encoder v2 (fz_enc2.cuh): the device source under the 32-thread warp model, or its sequential restatement"""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    out = np.empty(a.size + (a.size // SUB + 1) * 64 + 64, dtype=np.uint8)
    ns = C.c_uint64()
    n = _L.hm_encode_stream_v2(a.ctypes.data, a.size, out.ctypes.data, out.size, int(sequential), int(skip), C.byref(ns))
    assert n < 2**64 - 2, n
    return out[:n].copy(), int(ns.value)


def set_hist_sample(k: int):
    """0 / 1: the group's code is built from every sub-block; k > 1: from every k-th one (what the kernels do, k = 4), with
    a stand-in count of 1 for every symbol nobody counted."""
    _L.hm_set_hist_sample(C.c_int(k))


def encode_stream(a: np.ndarray):
    """a whole plane stream through the encoder the kernels run (device source, 32-thread warp model)"""
    return encode_stream_v2(a, sequential=False, skip=True)


def inflate(b: np.ndarray, n_out: int):
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty(n_out + 8, dtype=np.uint8)
    on, used = C.c_uint32(), C.c_uint64()
    rc = _L.hm_inflate(b.ctypes.data, b.size, out.ctypes.data, n_out, C.byref(on), C.byref(used))
    return rc, out[:on.value].copy(), int(used.value)


_L.hm_check_marker.restype = C.c_uint32
_L.hm_check_marker.argtypes = [C.c_void_p, C.c_uint32]
_L.hm_put_stored.restype = C.c_uint32
_L.hm_put_stored.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]


def check_marker(frag: np.ndarray) -> bool:
    frag = np.ascontiguousarray(frag, dtype=np.uint8)
    return bool(_L.hm_check_marker(frag.ctypes.data, frag.size))


def put_stored(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    out = np.empty(a.size + 32, dtype=np.uint8)
    n = _L.hm_put_stored(a.ctypes.data, a.size, out.ctypes.data)
    return out[:n].copy()


_L.hm_inflate_blockpar.restype = C.c_int
_L.hm_inflate_blockpar.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]


def inflate_blockpar(b: np.ndarray, n_out: int):
    """-> (rc, bytes, candidates found, candidates off the chain)"""
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros(n_out + 8, dtype=np.uint8)
    nc, nf = C.c_uint32(), C.c_uint32()
    rc = _L.hm_inflate_blockpar(b.ctypes.data, b.size, out.ctypes.data, n_out, C.byref(nc), C.byref(nf))
    return rc, out[:n_out].copy(), int(nc.value), int(nf.value)


def set_tile_pool_cap(n: int):
    """Capacity of the tile-record pool the measure pass fills for the write pass (0: the write pass searches again)."""
    _L.hm_set_tile_pool_cap(C.c_uint32(n))


def table_blocks() -> int:
    """Blocks the last inflate_blockpar call wrote from tile records."""
    _L.hm_get_table_blocks.restype = C.c_uint32
    return int(_L.hm_get_table_blocks())


_L.hm_lut_compare.restype = C.c_int
_L.hm_lut_compare.argtypes = [C.c_void_p]


def lut_compare(lens: np.ndarray) -> int:
    """entries that differ between the two ways of building the inflater's lookup table for the literal/length code with
    these 288 code lengths (-1: over-subscribed lengths)"""
    lens = np.ascontiguousarray(lens, dtype=np.uint8)
    assert lens.size == 288
    return int(_L.hm_lut_compare(lens.ctypes.data))
