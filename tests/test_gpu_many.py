"""GPU tests (-m gpu) of the list front end: many small files share passes of the kernels
(zip_compress_many / zip_uncompress_many, mzb_compress_host_many) and come out byte-identical to the one-file path."""
import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from conftest import synth_words

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = Path(__file__).resolve().parent.parent
CHK = 6 * 1048576


def _passes(L):
    a, b = C.c_uint64(), C.c_uint64()
    L.mzb_pass_counts(C.byref(a), C.byref(b))
    return a.value, b.value


def _many(L, fn, srcs, dsts, *extra):
    n = len(srcs)
    a = (C.c_char_p * n)(*[str(s).encode() for s in srcs])
    b = (C.c_char_p * n)(*[str(s).encode() for s in dsts])
    from datacompressionfloat_b200 import lib
    ctx = lib.CtxT()
    L.init_context(C.byref(ctx))
    rc = getattr(L, fn)(C.byref(ctx), n, a, b, *extra)
    assert rc == 0, rc
    return ctx


def test_many_small_files_share_passes_and_match_the_one_file_path(oracle):
    from datacompressionfloat_b200 import api, lib
    L = lib.load()
    bits = 9
    sizes = [CHK + 5000, 300, 0, 3 * CHK - 256, 70000, 2 * CHK, 1, CHK - 256 + 1, 9 * CHK, 123456] + [200000 + 1000 * i for i in range(54)]
    kinds = "GPSZR"
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(prefix="mrcz_many_", dir=base) as d:
        d = Path(d)
        srcs, words = [], []
        for i, n in enumerate(sizes):
            w = synth_words(kinds[i % 5], n, seed=i) if n else np.zeros(0, np.uint32)
            p = d / f"s{i:02d}.mrc"
            w.tofile(p)
            srcs.append(p); words.append(w)
        zm = [d / f"m{i:02d}.zip" for i in range(len(sizes))]
        z1 = [d / f"o{i:02d}.zip" for i in range(len(sizes))]
        p0 = _passes(L)
        ctx = _many(L, "zip_compress_many", srcs, zm, bits)
        p1 = _passes(L)
        assert ctx.fileCount == len(sizes)
        # 64 files, 63 of them small (one of 9 chunks goes alone, in 1 pass): 90-odd chunks in groups of <= 32
        assert p1[0] - p0[0] <= 8, p1[0] - p0[0]
        for s, z in zip(srcs, z1):
            api.zip_compress(str(s), str(z), bits)
        for a, b in zip(zm, z1):
            assert np.array_equal(np.fromfile(a, np.uint8), np.fromfile(b, np.uint8)), a.name
        keep = [i for i, n in enumerate(sizes) if n]      # an empty input has no container (workers.c:757-764)
        assert os.path.getsize(zm[2]) == 0
        outs = [d / f"u{i:02d}.mrc" for i in keep]
        p2 = _passes(L)
        _many(L, "zip_uncompress_many", [zm[i] for i in keep], outs)
        p3 = _passes(L)
        assert p3[1] - p2[1] <= 8, p3[1] - p2[1]
        for w, o in zip([words[i] for i in keep], outs):
            assert np.array_equal(np.fromfile(o, np.uint32), oracle.erasebytes(w.view(np.uint8), bits).view(np.uint32)), o.name


def test_many_items_api_against_single_calls(codec, oracle):
    from datacompressionfloat_b200 import lib
    L = lib.load()

    class ZipItem(C.Structure):
        _fields_ = [("h_words", C.c_void_p), ("nwords", C.c_uint64), ("exempt_words", C.c_uint32), ("reserved", C.c_uint32),
                    ("fsz", C.c_uint64), ("h_out", C.c_void_p), ("out_cap", C.c_size_t), ("out_size", C.c_uint64)]

    class UnzipItem(C.Structure):
        _fields_ = [("h_in", C.c_void_p), ("in_size", C.c_size_t), ("nwords", C.c_uint64), ("h_words_out", C.c_void_p),
                    ("out_cap_words", C.c_uint64)]
    chk, bits = 65536, 12
    ws = [synth_words(k, n, seed=n) for k, n in [("G", 65536 * 2), ("P", 100), ("S", 65536 * 3 - 256 + 17), ("Z", 40000), ("R", 65536 - 256)]]
    items = (ZipItem * len(ws))()
    outs = [np.zeros(codec.compress_bound(w.size, chk), np.uint8) for w in ws]
    for it, w, o in zip(items, ws, outs):
        it.h_words, it.nwords, it.exempt_words, it.fsz = w.ctypes.data, w.size, 256, w.size * 4
        it.h_out, it.out_cap = o.ctypes.data, o.size
    assert L.mzb_compress_host_many(codec._h, items, len(ws), bits, chk, 1) == 0
    for it, w, o in zip(items, ws, outs):
        one = codec.compress_host(w, bits, chk=chk)
        assert it.out_size == one.size and np.array_equal(o[:one.size], one)
    uitems = (UnzipItem * len(ws))()
    backs = [np.zeros(w.size, np.uint32) for w in ws]
    for ui, it, w, o, b in zip(uitems, items, ws, outs, backs):
        ui.h_in, ui.in_size, ui.nwords = o.ctypes.data + 17, it.out_size - 17, w.size
        ui.h_words_out, ui.out_cap_words = b.ctypes.data, b.size
    assert L.mzb_decompress_host_many(codec._h, uitems, len(ws), chk) == 0
    for w, b in zip(ws, backs):
        assert np.array_equal(b, oracle.erasebytes(w.view(np.uint8), bits).view(np.uint32))
    # too many chunks for one pass, odd chunk sizes: refused, not mangled
    assert L.mzb_compress_host_many(codec._h, items, len(ws), bits, 1000, 1) == lib.E_ARG
