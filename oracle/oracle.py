"""ctypes / subprocess front end to the CHECKER (oracle/liboracle.so and oracle/_ref/*).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_DIR = HERE / "_ref"
CHUNK_WORDS = 6 * 1048576
FILE_HEADER_BYTES = 17


def build(ref: bool = True) -> None:
    """Compile the restatement (and, when /root/reference exists, the reference itself)."""
    targets = ["all"] + (["ref"] if ref else [])
    # the reference's own mains linked against the product library (the drop-in proof): needs the built .so
    if ref and (HERE.parent / "datacompressionfloat_b200" / "libmrczip_b200.so").exists():
        targets.append("dropin")
    subprocess.run(["make", "-s", "-C", str(HERE)] + targets, check=True)


_lib = None


class _Stream(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("len", C.c_uint32), ("raw", C.c_uint32), ("n", C.c_uint32)]


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build(ref=False)
        L = C.CDLL(str(LIB_PATH))
        u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
        L.orc_mask_for_bits.restype = C.c_uint32
        L.orc_mask_for_bits.argtypes = [C.c_int]
        L.orc_apply_mask.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int]
        L.orc_split.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_int, C.c_int]
        L.orc_merge.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.orc_pack_header.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.orc_unpack_header.argtypes = [C.c_void_p, C.POINTER(C.c_int), u32p]
        L.orc_erasebytes.restype = C.c_int64
        L.orc_erasebytes.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.orc_compress_bound.restype = C.c_size_t
        L.orc_compress_bound.argtypes = [C.c_uint64, C.c_uint32]
        L.orc_compress.restype = C.c_int64
        L.orc_compress.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_decompress.restype = C.c_int64
        L.orc_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_decompress_noz.restype = C.c_int64
        L.orc_decompress_noz.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_inflate_raw.restype = C.c_int
        L.orc_inflate_raw.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.orc_parse_container.restype = C.c_int64
        L.orc_parse_container.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), u32p,
                                          C.POINTER(_Stream), C.c_size_t]
        L.orc_zlib_version.restype = C.c_char_p
        _lib = L
    return _lib


def _bytes_arr(b) -> np.ndarray:
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b.view(np.uint8).reshape(-1)
    return np.ascontiguousarray(a)


def mask_for_bits(bits: int) -> int:
    return int(lib().orc_mask_for_bits(bits))


def split(words: np.ndarray, bits: int, is_first_chunk: bool):
    """Returns (masked_words, [plane0..plane3]) for ONE chunk (workers.c:180-203)."""
    w = np.array(words, dtype=np.uint32, copy=True).reshape(-1)
    n = w.size
    planes = [np.empty(n, dtype=np.uint8) for _ in range(4)]
    ptrs = (C.c_void_p * 4)(*[p.ctypes.data for p in planes])
    lib().orc_split(w.ctypes.data, n, ptrs, bits, int(is_first_chunk))
    return w, planes


def split_file(words: np.ndarray, bits: int, chk: int = CHUNK_WORDS):
    """Chunk loop of run_compress: first chunk of the FILE keeps its first 256 words unmasked."""
    w = np.asarray(words, dtype=np.uint32).reshape(-1)
    outs = [[], [], [], []]
    masked = []
    for k, w0 in enumerate(range(0, w.size, chk)):
        m, pl = split(w[w0:w0 + chk], bits, k == 0)
        masked.append(m)
        for j in range(4):
            outs[j].append(pl[j])
    if not masked:
        return w.copy(), [np.empty(0, np.uint8) for _ in range(4)]
    return np.concatenate(masked), [np.concatenate(o) for o in outs]


def merge(planes) -> np.ndarray:
    planes = [np.ascontiguousarray(p, dtype=np.uint8) for p in planes]
    n = planes[0].size
    out = np.empty(n, dtype=np.uint32)
    ptrs = (C.c_void_p * 4)(*[p.ctypes.data for p in planes])
    lib().orc_merge(out.ctypes.data, n, ptrs)
    return out


def erasebytes(file_bytes, bits: int) -> np.ndarray:
    a = _bytes_arr(file_bytes)
    out = np.empty(a.size, dtype=np.uint8)
    n = lib().orc_erasebytes(a.ctypes.data, a.size, bits, out.ctypes.data)
    return out[:n]


def compress(file_bytes, bits: int, chk: int = CHUNK_WORDS, reset_per_chunk: bool = False) -> np.ndarray:
    a = _bytes_arr(file_bytes)
    cap = lib().orc_compress_bound(a.size, chk)
    out = np.empty(cap, dtype=np.uint8)
    n = lib().orc_compress(a.ctypes.data, a.size, bits, chk, out.ctypes.data, cap, int(reset_per_chunk))
    if n < 0:
        raise RuntimeError(f"orc_compress failed: {n}")
    return out[:n].copy()


def parse_container(container):
    a = _bytes_arr(container)
    fsz, chk = C.c_uint64(), C.c_uint32()
    ns = lib().orc_parse_container(a.ctypes.data, a.size, C.byref(fsz), C.byref(chk), None, 0)
    if ns < 0:
        raise RuntimeError(f"orc_parse_container failed: {ns}")
    arr = (_Stream * max(ns, 1))()
    lib().orc_parse_container(a.ctypes.data, a.size, C.byref(fsz), C.byref(chk), arr, ns)
    streams = [dict(offset=arr[i].offset, len=arr[i].len, raw=bool(arr[i].raw), n=arr[i].n) for i in range(ns)]
    return fsz.value, chk.value, streams


def decompress(container, use_zlib: bool = True) -> np.ndarray:
    a = _bytes_arr(container)
    if a.size < FILE_HEADER_BYTES:
        return np.empty(0, dtype=np.uint8)
    fsz = int(a[:8].view(np.uint64)[0])
    out = np.empty((fsz // 4) * 4, dtype=np.uint8)
    fn = lib().orc_decompress if use_zlib else lib().orc_decompress_noz
    n = fn(a.ctypes.data, a.size, out.ctypes.data, out.size)
    if n < 0:
        raise RuntimeError(f"orc_decompress failed: {n}")
    return out[:n]


def inflate_raw(payload, n_out: int):
    """Independent (no zlib) raw inflate of one payload; returns (bytes, consumed)."""
    a = _bytes_arr(payload)
    out = np.empty(max(n_out, 1), dtype=np.uint8)
    on, used = C.c_size_t(), C.c_size_t()
    rc = lib().orc_inflate_raw(a.ctypes.data, a.size, out.ctypes.data, n_out, C.byref(on), C.byref(used))
    if rc != 0:
        raise RuntimeError(f"orc_inflate_raw failed: {rc}")
    return out[:on.value], used.value


def zlib_version() -> str:
    return lib().orc_zlib_version().decode()


# ----------------------------------------------------------------------------- the real reference (oracle/_ref)

def have_ref() -> bool:
    return all((REF_DIR / b).exists() for b in ("mrc_tar_c", "mrc_tarx_c", "erasebytes_c", "ref_harness"))


def _run(args, **kw):
    return subprocess.run([str(a) for a in args], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


def _tmpdir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.TemporaryDirectory(prefix="orc_", dir=base)


def ref_compress(file_bytes, bits: int) -> np.ndarray:
    """mrc_tar_c -t zip (reference binary, its own zlib 1.2.8)."""
    with _tmpdir() as d:
        src, dst = Path(d) / "x.mrc", Path(d) / "x.zip"
        _bytes_arr(file_bytes).tofile(src)
        _run([REF_DIR / "mrc_tar_c", "-i", src, "-o", dst, "-b", bits, "-t", "zip"])
        return np.fromfile(dst, dtype=np.uint8)


def ref_decompress(container) -> np.ndarray:
    """mrc_tar_c -t unzip (reference binary, its own zlib 1.2.8)."""
    with _tmpdir() as d:
        src, dst = Path(d) / "x.zip", Path(d) / "x.out"
        _bytes_arr(container).tofile(src)
        _run([REF_DIR / "mrc_tar_c", "-i", src, "-o", dst, "-t", "unzip"])
        return np.fromfile(dst, dtype=np.uint8)


def ref_erasebytes(file_bytes, bits: int) -> np.ndarray:
    with _tmpdir() as d:
        src, dst = Path(d) / "x.mrc", Path(d) / "x.out"
        _bytes_arr(file_bytes).tofile(src)
        _run([REF_DIR / "erasebytes_c", "-i", src, "-o", dst, "-b", bits])
        return np.fromfile(dst, dtype=np.uint8)


def ref_split(file_bytes, bits: int):
    """Reference split_float_to_byte_stream over the run_compress chunk loop -> (masked words, 4 planes)."""
    with _tmpdir() as d:
        src, pre = Path(d) / "x.mrc", Path(d) / "o"
        _bytes_arr(file_bytes).tofile(src)
        _run([REF_DIR / "ref_harness", "split", src, bits, pre])
        masked = np.fromfile(f"{pre}.masked", dtype=np.uint32)
        return masked, [np.fromfile(f"{pre}.p{j}", dtype=np.uint8) for j in range(4)]


def ref_merge(planes) -> np.ndarray:
    with _tmpdir() as d:
        names = []
        for j, p in enumerate(planes):
            names.append(Path(d) / f"p{j}")
            np.ascontiguousarray(p, dtype=np.uint8).tofile(names[-1])
        out = Path(d) / "m"
        _run([REF_DIR / "ref_harness", "merge", *names, out])
        return np.fromfile(out, dtype=np.uint32)


def ref_inflate(payload, n_out: int) -> np.ndarray:
    """Reference mzlib_inf (zip.c:262) with the reference's libz 1.2.8 on one payload."""
    with _tmpdir() as d:
        src, dst = Path(d) / "p.bin", Path(d) / "p.out"
        _bytes_arr(payload).tofile(src)
        _run([REF_DIR / "ref_harness", "inflate", src, n_out, dst])
        return np.fromfile(dst, dtype=np.uint8)


def ref_zlib_version() -> str:
    return _run([REF_DIR / "ref_harness", "version"]).stdout.decode().strip()
