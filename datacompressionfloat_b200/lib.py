"""ctypes binding of libmrczip_b200.so (the C ABI declared in include/mrczip_b200.h).

There is no Python or CPU fallback: if the shared library is missing this module raises, and every
compute call needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libmrczip_b200.so"
CSRC_DIR = PKG_DIR / "csrc"

OK, E_ARG, E_CUDA, E_NOMEM, E_FORMAT, E_SPACE, E_IO = 0, -1, -2, -3, -4, -5, -6
CHUNK_WORDS = 6 * 1048576
FILE_HEADER_BYTES = 17
MRC_HEADER_WORDS = 256
SUB_BYTES = 16384


class MzbError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(f"{what}: {strerror(code)} ({code})")


class Stats(C.Structure):
    _fields_ = [("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64), ("chunks", C.c_uint32), ("streams", C.c_uint32),
                ("raw_streams", C.c_uint32), ("stored_subblocks", C.c_uint32), ("general_streams", C.c_uint32),
                ("fast_failed", C.c_uint32), ("kernel_launches", C.c_uint32), ("blockpar_streams", C.c_uint32),
                ("zero_subblocks", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k not in ("pad", "reserved")}


class CtxT(C.Structure):
    """ctx_t, reference src/include/common.h:33-41"""
    _fields_ = [("fileCount", C.c_uint32), ("allFileSize", C.c_uint64), ("allZipFileSize", C.c_uint64),
                ("zipTime", C.c_double), ("unzipTime", C.c_double)]


class MrczipHeaderT(C.Structure):
    """mrczip_header_t, reference src/include/common.h:50-56"""
    _fields_ = [("fsz", C.c_uint64), ("chk", C.c_uint32), ("type", C.c_char), ("ztypes", C.c_char * 4)]


# every symbol include/mrczip_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "run_compress", "run_uncompress", "zip_compress", "zip_uncompress", "zip_compress_many", "zip_uncompress_many", "pack_header", "unpack_header",
    "init_context", "reset_context", "update_context", "print_context_info", "init_mrczip_header",
    "read_mrczip_header", "write_mrczip_header", "print_mrczip_header", "get_file_size", "now_sec",
    "isTestThroughput",
    "mzb_create", "mzb_create_on_stream", "mzb_destroy", "mzb_set_io_threads", "mzb_pass_counts", "mzb_compress_host_many", "mzb_decompress_host_many", "mzb_device_count", "mzb_set_devices", "mzb_mrc_parse", "mzb_set_mrc_aware", "mzb_error_report_device", "mzb_error_report_host", "mzb_set_batch_chunks", "mzb_set_variant", "mzb_set_inflate_variant", "mzb_compress_bound",
    "mzb_compress_device", "mzb_decompress_device", "mzb_mask_split_device", "mzb_merge_device",
    "mzb_compress_host", "mzb_decompress_host", "mzb_host_alloc", "mzb_host_free", "mzb_last_stats",
    "mzb_set_profiling", "mzb_stage_count", "mzb_stage_name", "mzb_stage_ms",
    "mzb_version", "mzb_strerror",
]


def build(verbose: bool = False) -> Path:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(CSRC_DIR)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("building libmrczip_b200.so failed")
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, u64, u32, i32, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t
    L.mzb_create.restype = i32
    L.mzb_create.argtypes = [C.POINTER(vp), i32, vp]
    L.mzb_create_on_stream.restype = i32
    L.mzb_create_on_stream.argtypes = [C.POINTER(vp), i32, vp]
    L.mzb_destroy.restype = None
    L.mzb_destroy.argtypes = [vp]
    L.mzb_set_batch_chunks.restype = i32
    L.mzb_set_batch_chunks.argtypes = [vp, u32]
    L.mzb_set_variant.restype = i32
    L.mzb_set_variant.argtypes = [vp, i32, i32]
    L.mzb_set_inflate_variant.restype = i32
    L.mzb_set_inflate_variant.argtypes = [vp, i32]
    L.mzb_compress_bound.restype = sz
    L.mzb_compress_bound.argtypes = [u64, u32]
    L.mzb_compress_device.restype = i32
    L.mzb_compress_device.argtypes = [vp, vp, u64, i32, u32, u32, u64, i32, vp, sz, C.POINTER(u64)]
    L.mzb_decompress_device.restype = i32
    L.mzb_decompress_device.argtypes = [vp, vp, sz, i32, u32, u64, vp, u64, C.POINTER(u64)]
    L.mzb_mask_split_device.restype = i32
    L.mzb_mask_split_device.argtypes = [vp, vp, u64, i32, u32, vp, u64]
    L.mzb_merge_device.restype = i32
    L.mzb_merge_device.argtypes = [vp, vp, u64, u64, vp]
    L.mzb_compress_host.restype = i32
    L.mzb_compress_host.argtypes = [vp, vp, u64, i32, u32, u32, u64, i32, vp, sz, C.POINTER(u64)]
    L.mzb_decompress_host.restype = i32
    L.mzb_decompress_host.argtypes = [vp, vp, sz, i32, u32, u64, vp, u64, C.POINTER(u64)]
    L.mzb_host_alloc.restype = vp
    L.mzb_host_alloc.argtypes = [sz]
    L.mzb_host_free.restype = None
    L.mzb_host_free.argtypes = [vp]
    L.mzb_last_stats.restype = i32
    L.mzb_last_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mzb_set_profiling.restype = i32
    L.mzb_set_profiling.argtypes = [vp, i32]
    L.mzb_stage_count.restype = i32
    L.mzb_stage_name.restype = C.c_char_p
    L.mzb_stage_name.argtypes = [i32]
    L.mzb_stage_ms.restype = i32
    L.mzb_stage_ms.argtypes = [vp, C.POINTER(C.c_float), i32]
    L.mzb_version.restype = C.c_char_p
    L.mzb_strerror.restype = C.c_char_p
    L.mzb_strerror.argtypes = [i32]
    L.zip_compress.restype = i32
    L.zip_compress.argtypes = [C.POINTER(CtxT), C.c_char_p, C.c_char_p, i32]
    L.zip_uncompress.restype = i32
    L.zip_uncompress.argtypes = [C.POINTER(CtxT), C.c_char_p, C.c_char_p]
    L.pack_header.restype = None
    L.pack_header.argtypes = [C.c_char_p, i32, u32]
    L.unpack_header.restype = None
    L.unpack_header.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(u32)]
    L.init_context.restype = None
    L.init_context.argtypes = [C.POINTER(CtxT)]
    _lib = L
    return L


def strerror(code: int) -> str:
    return load().mzb_strerror(code).decode()


def check(code: int, what: str) -> None:
    if code != OK:
        raise MzbError(code, what)
