"""Host-side mirror of the reference interface for the float32 compress / decompress path.

`Codec` owns one mzb_ctx (one CUDA device + stream) and exposes, on torch device tensors or numpy /
pinned host buffers, the operations the reference performs inside run_compress / run_uncompress
(reference src/core/workers.c:690-881, 568-688):

    mask_split   apply_mask + split_float_to_byte_stream   (workers.c:82-101, 180-203)
    merge        merge_byte_to_float_stream                (workers.c:423-442)
    compress     chunk loop + per-plane deflate + container (workers.c:779-855, zip.c:164-196)
    decompress   chunk loop + per-plane inflate + merge     (workers.c:592-672, zip.c:262-284)

and the path-level calls zip_compress / zip_uncompress (reference src/core/adapt.c:28-90), which go
through the library's C host layer.  PyTorch is used for device memory and streams only; all compute
is in libmrczip_b200.so.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import lib as _lib
from .lib import CHUNK_WORDS, FILE_HEADER_BYTES, MRC_HEADER_WORDS, MzbError, check

try:  # torch is plumbing (device memory / streams), not a hard requirement of the C ABI
    import torch
except Exception:  # pragma: no cover
    torch = None


def _need_torch():
    if torch is None:
        raise ImportError("torch is required for device-tensor calls")


class Codec:
    """One GPU context.  Not thread-safe: use one Codec per thread (like one reference worker thread)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None, batch_chunks: Optional[int] = None):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.device = int(device)
        self._private_stream = stream is None
        if stream is None:   # the context makes its own non-blocking stream
            check(self._L.mzb_create(C.byref(self._h), self.device, None), "mzb_create")
        else:                # a cudaStream_t handle; 0 is the legacy default stream
            check(self._L.mzb_create_on_stream(C.byref(self._h), self.device, C.c_void_p(int(stream))), "mzb_create_on_stream")
        if batch_chunks:
            check(self._L.mzb_set_batch_chunks(self._h, int(batch_chunks)), "mzb_set_batch_chunks")

    @classmethod
    def on_current_stream(cls, batch_chunks: Optional[int] = None) -> "Codec":
        """Context bound to torch's current device and current stream (so torch.cuda.Event timing sees the kernels)."""
        _need_torch()
        dev = torch.cuda.current_device()
        return cls(dev, torch.cuda.current_stream(dev).cuda_stream, batch_chunks)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.mzb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ knobs / info
    def set_variant(self, split: int = 0, merge: int = 0):
        check(self._L.mzb_set_variant(self._h, split, merge), "mzb_set_variant")

    def set_inflate_variant(self, v: int = 0):
        """0: lean table-loop inflater first (default); 1: the full group inflater alone (its fallback)."""
        check(self._L.mzb_set_inflate_variant(self._h, v), "mzb_set_inflate_variant")

    def set_batch_chunks(self, n: int):
        check(self._L.mzb_set_batch_chunks(self._h, n), "mzb_set_batch_chunks")

    def set_profiling(self, on: bool = True):
        check(self._L.mzb_set_profiling(self._h, int(on)), "mzb_set_profiling")

    def stage_ms(self) -> dict:
        """Device time per pipeline stage of the last call (CUDA events on the context's stream)."""
        n = self._L.mzb_stage_count()
        buf = (C.c_float * n)()
        check(self._L.mzb_stage_ms(self._h, buf, n), "mzb_stage_ms")
        return {self._L.mzb_stage_name(i).decode(): float(buf[i]) for i in range(n)}

    def stats(self) -> dict:
        s = _lib.Stats()
        check(self._L.mzb_last_stats(self._h, C.byref(s)), "mzb_last_stats")
        return s.as_dict()

    @staticmethod
    def compress_bound(nwords: int, chk: int = CHUNK_WORDS) -> int:
        return int(_lib.load().mzb_compress_bound(nwords, chk))

    # ------------------------------------------------------------------ device tensors
    def _order_inputs(self):
        """A context with a private stream is not ordered after torch's current stream: wait for the producers of
        the tensors we are about to read (the library synchronises its own stream before returning)."""
        if self._private_stream:
            torch.cuda.current_stream(self.device).synchronize()

    def _dev_words(self, words):
        _need_torch()
        if not (words.is_cuda and words.is_contiguous() and words.element_size() == 4):
            raise ValueError("words must be a contiguous 4-byte-element CUDA tensor")
        self._order_inputs()
        return words

    def mask_split(self, words, bits: int, exempt_words: int = MRC_HEADER_WORDS, out=None):
        """-> uint8 tensor [4, stride]; plane j = out[j, :nwords] (bit-identical to the reference's zins[j])."""
        w = self._dev_words(words)
        n = w.numel()
        stride = (n + 15) // 16 * 16
        if out is None:
            out = torch.empty((4, max(stride, 16)), dtype=torch.uint8, device=w.device)
        stride = out.stride(0)
        check(self._L.mzb_mask_split_device(self._h, w.data_ptr(), n, bits, exempt_words, out.data_ptr(), stride),
              "mzb_mask_split_device")
        return out

    def merge(self, planes, nwords: int, out=None):
        """planes: uint8 tensor [4, stride] -> int32 tensor [nwords]."""
        _need_torch()
        self._order_inputs()
        if out is None:
            out = torch.empty(max(nwords, 4), dtype=torch.int32, device=planes.device)
        check(self._L.mzb_merge_device(self._h, planes.data_ptr(), planes.stride(0), nwords, out.data_ptr()),
              "mzb_merge_device")
        return out[:nwords]

    def compress(self, words, bits: int, chk: int = CHUNK_WORDS, fsz: Optional[int] = None,
                 exempt_words: int = MRC_HEADER_WORDS, write_file_header: bool = True, out=None):
        """Device-resident compress; returns a uint8 CUDA tensor view holding the container (or chunk records)."""
        w = self._dev_words(words)
        n = w.numel()
        cap = self.compress_bound(n, chk)
        if out is None:
            out = torch.empty(cap, dtype=torch.uint8, device=w.device)
        size = C.c_uint64()
        check(self._L.mzb_compress_device(self._h, w.data_ptr(), n, bits, exempt_words, chk,
                                          n * 4 if fsz is None else fsz, int(write_file_header), out.data_ptr(),
                                          out.numel(), C.byref(size)), "mzb_compress_device")
        return out[:size.value]

    def decompress(self, container, has_file_header: bool = True, chk: int = CHUNK_WORDS, nwords: int = 0, out=None):
        """Device-resident decompress; returns an int32 CUDA tensor [nwords]."""
        _need_torch()
        if not (container.is_cuda and container.is_contiguous() and container.dtype == torch.uint8):
            raise ValueError("container must be a contiguous uint8 CUDA tensor")
        self._order_inputs()
        if has_file_header:
            if container.numel() == 0:
                return torch.empty(0, dtype=torch.int32, device=container.device)
            if container.numel() < FILE_HEADER_BYTES:
                raise MzbError(_lib.E_FORMAT, "decompress")
            hdr = container[:FILE_HEADER_BYTES].cpu().numpy()
            nwords = int(hdr[:8].view(np.uint64)[0]) // 4
        if out is None:
            out = torch.empty(max(nwords, 4), dtype=torch.int32, device=container.device)
        got = C.c_uint64()
        check(self._L.mzb_decompress_device(self._h, container.data_ptr(), container.numel(), int(has_file_header), chk,
                                            nwords, out.data_ptr(), out.numel(), C.byref(got)), "mzb_decompress_device")
        return out[:got.value]

    # ------------------------------------------------------------------ host buffers (end to end)
    def compress_host(self, words: np.ndarray, bits: int, chk: int = CHUNK_WORDS, fsz: Optional[int] = None,
                      exempt_words: int = MRC_HEADER_WORDS, write_file_header: bool = True, out: Optional[np.ndarray] = None):
        w = np.ascontiguousarray(words).view(np.uint32).reshape(-1)
        cap = self.compress_bound(w.size, chk)
        if out is None:
            out = np.empty(cap, dtype=np.uint8)
        size = C.c_uint64()
        check(self._L.mzb_compress_host(self._h, w.ctypes.data, w.size, bits, exempt_words, chk,
                                        w.size * 4 if fsz is None else fsz, int(write_file_header), out.ctypes.data,
                                        out.size, C.byref(size)), "mzb_compress_host")
        return out[:size.value]

    def decompress_host(self, container: np.ndarray, has_file_header: bool = True, chk: int = CHUNK_WORDS,
                        nwords: int = 0, out: Optional[np.ndarray] = None):
        c = np.ascontiguousarray(container).view(np.uint8).reshape(-1)
        if has_file_header:
            if c.size == 0:
                return np.empty(0, dtype=np.uint32)
            if c.size < FILE_HEADER_BYTES:
                raise MzbError(_lib.E_FORMAT, "decompress_host")
            nwords = int(c[:8].view(np.uint64)[0]) // 4
        if out is None:
            out = np.empty(max(nwords, 1), dtype=np.uint32)
        got = C.c_uint64()
        check(self._L.mzb_decompress_host(self._h, c.ctypes.data, c.size, int(has_file_header), chk, nwords,
                                          out.ctypes.data, out.size, C.byref(got)), "mzb_decompress_host")
        return out[:got.value]

    def compress_host_ptr(self, in_ptr: int, nwords: int, bits: int, out_ptr: int, out_cap: int, chk: int = CHUNK_WORDS,
                          fsz: Optional[int] = None, exempt_words: int = MRC_HEADER_WORDS, write_file_header: bool = True) -> int:
        """Raw-pointer form for pinned buffers (bench.py's end-to-end leg)."""
        size = C.c_uint64()
        check(self._L.mzb_compress_host(self._h, in_ptr, nwords, bits, exempt_words, chk,
                                        nwords * 4 if fsz is None else fsz, int(write_file_header), out_ptr, out_cap,
                                        C.byref(size)), "mzb_compress_host")
        return int(size.value)

    def decompress_host_ptr(self, in_ptr: int, in_size: int, out_ptr: int, out_cap_words: int, has_file_header: bool = True,
                            chk: int = CHUNK_WORDS, nwords: int = 0) -> int:
        got = C.c_uint64()
        check(self._L.mzb_decompress_host(self._h, in_ptr, in_size, int(has_file_header), chk, nwords, out_ptr,
                                          out_cap_words, C.byref(got)), "mzb_decompress_host")
        return int(got.value)


# ---------------------------------------------------------------------- path-level calls (reference adapt.c:28-90)

def zip_compress(src: str, dst: str, bits_to_loss: int = 0) -> dict:
    """zip_compress(ctx, src, dst, bitsToLoss) through the library's C host layer; returns the ctx_t fields."""
    L = _lib.load()
    ctx = _lib.CtxT()
    L.init_context(C.byref(ctx))
    check(L.zip_compress(C.byref(ctx), src.encode(), dst.encode(), int(bits_to_loss)), "zip_compress")
    return {k: getattr(ctx, k) for k, _ in ctx._fields_}


def zip_uncompress(src: str, dst: str) -> dict:
    L = _lib.load()
    ctx = _lib.CtxT()
    L.init_context(C.byref(ctx))
    check(L.zip_uncompress(C.byref(ctx), src.encode(), dst.encode()), "zip_uncompress")
    return {k: getattr(ctx, k) for k, _ in ctx._fields_}


# ---------------------------------------------------------------------- container helpers (host, tiny)

def file_header(fsz: int, chk: int = CHUNK_WORDS) -> np.ndarray:
    """The 17-byte file header (reference common.c:137-149): u64 fsz, u32 chk, u8 type, u8 ztypes[4]."""
    h = np.zeros(FILE_HEADER_BYTES, dtype=np.uint8)
    h[:8] = np.frombuffer(np.uint64(fsz).tobytes(), dtype=np.uint8)
    h[8:12] = np.frombuffer(np.uint32(chk).tobytes(), dtype=np.uint8)
    return h


def chunk_range(nchunks: int, rank: int, world: int):
    """Contiguous chunk range of `rank`: [rank * ceil(C/G), min(C, (rank+1) * ceil(C/G)))."""
    per = -(-nchunks // world)
    lo = min(nchunks, rank * per)
    return lo, min(nchunks, lo + per)


def segment_offsets(segment_sizes):
    """Exclusive scan of the per-rank segment byte counts -> each segment's offset in the container."""
    offs, run = [], FILE_HEADER_BYTES
    for s in segment_sizes:
        offs.append(run)
        run += int(s)
    return offs, run
