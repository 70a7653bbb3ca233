"""GPU tests (-m gpu) of SURVEY 8f #3: MRC-aware compression (extended header, non-float modes) and the error report
(reference src/tool/erroranalysis.c:188-220) computed on the GPU, against a numpy restatement and against the
reference's own erroranalysis_c."""
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
REF = ROOT / "oracle" / "_ref"


class ErrRep(C.Structure):
    _fields_ = [("count", C.c_uint64), ("nan_count", C.c_uint64), ("max_abs_err", C.c_float), ("max_abs_n1", C.c_float),
                ("max_abs_n2", C.c_float), ("max_rel_err", C.c_float), ("max_rel_n1", C.c_float), ("max_rel_n2", C.c_float),
                ("max_abs_index", C.c_uint64), ("max_rel_index", C.c_uint64), ("sum_abs_err", C.c_double)]


def np_report(a: np.ndarray, b: np.ndarray):
    """erroranalysis.c:188-220 in numpy (float32 arithmetic)"""
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        err = np.abs(b - a).astype(np.float32)
        rel = np.where(np.abs(a) > np.float32(10E-4), err / np.abs(a), np.float32(0)).astype(np.float32)
    ok = ~np.isnan(err)
    e2 = np.where(ok, err, np.float32(-1))
    r2 = np.where(ok & ~np.isnan(rel), rel, np.float32(-1))
    ia, ir = int(np.argmax(e2)), int(np.argmax(r2))
    return dict(nan=int((~ok).sum()), max_abs=float(e2[ia]), ia=ia, max_rel=float(r2[ir]), ir=ir,
                s=float(err[ok].astype(np.float64).sum()))


def _report(codec, fn, *args):
    L = codec._L
    f = getattr(L, fn)
    f.restype = C.c_int
    out = ErrRep()
    rc = f(codec._h, *args, C.byref(out))
    assert rc == 0, rc
    return out


@pytest.mark.parametrize("bits", [0, 5, 12, 23, 32])
def test_error_report_of_the_mask_matches_numpy(codec, bits):
    w = synth_words("G", 1_000_003, seed=bits)
    a = w.view(np.float32)
    mask = np.uint32((0xFFFFFFFF << bits) & 0xFFFFFFFF if bits < 32 else 0)
    m = w.copy()
    m[256:] &= mask
    want = np_report(a, m.view(np.float32))
    d = torch.from_numpy(w.view(np.int32)).cuda()
    r = _report(codec, "mzb_error_report_device", C.c_void_p(d.data_ptr()), C.c_void_p(0), C.c_uint64(w.size), C.c_int(bits), C.c_uint32(256))
    assert r.count == w.size and r.nan_count == want["nan"]
    assert r.max_abs_err == np.float32(want["max_abs"]) and r.max_abs_index == want["ia"]
    assert r.max_rel_err == np.float32(want["max_rel"]) and r.max_rel_index == want["ir"]
    assert r.max_abs_n1 == a[want["ia"]] and r.max_abs_n2 == m.view(np.float32)[want["ia"]]
    assert abs(r.sum_abs_err - want["s"]) <= 1e-4 * max(want["s"], 1e-30)
    # two-buffer form, device and host, on the masked copy
    d2 = torch.from_numpy(m.view(np.int32)).cuda()
    r2 = _report(codec, "mzb_error_report_device", C.c_void_p(d.data_ptr()), C.c_void_p(d2.data_ptr()), C.c_uint64(w.size), C.c_int(0), C.c_uint32(0))
    r3 = _report(codec, "mzb_error_report_host", C.c_void_p(w.ctypes.data), C.c_void_p(m.ctypes.data), C.c_uint64(w.size), C.c_int(0), C.c_uint32(0))
    for x in (r2, r3):
        assert (x.max_abs_err, x.max_abs_index, x.max_rel_err, x.max_rel_index, x.nan_count) == \
               (r.max_abs_err, r.max_abs_index, r.max_rel_err, r.max_rel_index, r.nan_count)


def test_error_report_special_values(codec):
    a = np.array([1.0, np.nan, np.inf, -np.inf, 0.0, 1e-4, -3.5, 2.0], dtype=np.float32)
    b = np.array([1.5, 1.0, np.inf, np.inf, 0.25, 2e-4, np.nan, 2.0], dtype=np.float32)
    want = np_report(a, b)
    r = _report(codec, "mzb_error_report_host", C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_uint64(a.size), C.c_int(0), C.c_uint32(0))
    assert r.nan_count == want["nan"] == 3                       # nan - x, inf - inf, x - nan
    assert r.max_abs_err == np.float32(np.inf) and r.max_abs_index == 3
    assert r.max_rel_index == want["ir"]
    e = _report(codec, "mzb_error_report_host", C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_uint64(0), C.c_int(0), C.c_uint32(0))
    assert e.count == 0 and e.max_abs_index == 2**64 - 1


def test_error_report_next_to_the_reference_tool(codec, oracle):
    """erroranalysis_c -a orig -b roundtrip -k 1 (what run_full_test.sh:106 prints): its top point is ours"""
    if not (REF / "erroranalysis_c").exists():
        pytest.skip("oracle/_ref/erroranalysis_c not present")
    w = synth_words("G", 300_000, seed=5)
    w[:256] = 0                                                     # the tool reads the header words as floats too
    bits = 14
    cont = codec.compress(torch.from_numpy(w.view(np.int32)).cuda(), bits)
    back = codec.decompress(cont).cpu().numpy().view(np.uint32)
    with tempfile.TemporaryDirectory() as d:
        pa, pb = Path(d) / "a.mrc", Path(d) / "b.mrc"
        w.tofile(pa); back.tofile(pb)
        out = subprocess.run([str(REF / "erroranalysis_c"), "-a", str(pa), "-b", str(pb), "-k", "1"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, check=True).stdout.decode().split()
    n1, n2, err = float(out[0]), float(out[1]), float(out[3])
    r = _report(codec, "mzb_error_report_host", C.c_void_p(w.ctypes.data), C.c_void_p(back.ctypes.data), C.c_uint64(w.size), C.c_int(0), C.c_uint32(0))
    assert abs(r.max_abs_err - err) <= 1e-6 * err                   # the tool prints %E (6 digits)
    assert abs(r.max_abs_n1 - n1) <= 1e-5 * abs(n1) + 1e-6 and abs(r.max_abs_n2 - n2) <= 1e-5 * abs(n2) + 1e-6


def _mrc_file(path, mode, next_bytes, data_words, seed):
    rng = np.random.default_rng(seed)
    h = np.zeros(256, dtype=np.int32)
    h[0:3] = (data_words, 1, 1)
    h[3] = mode
    h[23] = next_bytes
    ext = rng.integers(0, 2**32, next_bytes // 4, dtype=np.uint64).astype(np.uint32)
    data = rng.standard_normal(data_words, dtype=np.float32).view(np.uint32)
    w = np.concatenate([h.view(np.uint32), ext, data])
    w.tofile(path)
    return w


def test_mrc_aware_compression_extended_header_and_modes(codec, oracle):
    from datacompressionfloat_b200 import api, lib
    L = lib.load()
    L.mzb_set_mrc_aware.argtypes = [C.c_int]
    bits = 10
    mask = np.uint32((0xFFFFFFFF << bits) & 0xFFFFFFFF)
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        try:
            # (a) extended header of 4096 bytes: by default (reference behaviour) its words are masked ...
            w = _mrc_file(d / "x.mrc", 2, 4096, 200_000, 1)
            api.zip_compress(str(d / "x.mrc"), str(d / "x.zip"), bits)
            api.zip_uncompress(str(d / "x.zip"), str(d / "x.out"))
            got = np.fromfile(d / "x.out", dtype=np.uint32)
            assert np.array_equal(got, oracle.erasebytes(w.view(np.uint8), bits).view(np.uint32))
            # ... MRC-aware they are not, the data behind them still is
            L.mzb_set_mrc_aware(1)
            api.zip_compress(str(d / "x.mrc"), str(d / "x2.zip"), bits)
            api.zip_uncompress(str(d / "x2.zip"), str(d / "x2.out"))
            got = np.fromfile(d / "x2.out", dtype=np.uint32)
            exp = w.copy()
            exp[256 + 1024:] &= mask
            assert np.array_equal(got, exp)
            # (b) mode 1 (int16): nothing is erased
            w = _mrc_file(d / "i.mrc", 1, 0, 100_000, 2)
            api.zip_compress(str(d / "i.mrc"), str(d / "i.zip"), bits)
            api.zip_uncompress(str(d / "i.zip"), str(d / "i.out"))
            assert np.array_equal(np.fromfile(d / "i.out", dtype=np.uint32), w)
            # (c) not an MRC header at all: the reference's behaviour
            w = synth_words("R", 50_000, seed=3)
            w[0] = 0
            w.tofile(d / "r.bin")
            api.zip_compress(str(d / "r.bin"), str(d / "r.zip"), bits)
            api.zip_uncompress(str(d / "r.zip"), str(d / "r.out"))
            assert np.array_equal(np.fromfile(d / "r.out", dtype=np.uint32), oracle.erasebytes(w.view(np.uint8), bits).view(np.uint32))
        finally:
            L.mzb_set_mrc_aware(0)
