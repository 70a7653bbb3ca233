"""GPU test (-m gpu, needs 2 devices; skipped on a one-GPU box): one file cut over two devices behind the reference's
own entry points (zip_compress / zip_uncompress) gives the byte-identical container and the same bits back."""
import ctypes as C
import os
import tempfile
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_one_file_on_two_devices_is_byte_identical(oracle):
    from datacompressionfloat_b200 import api, lib, synth
    L = lib.load()
    if L.mzb_device_count() < 2:
        pytest.skip("needs two CUDA devices (run with gpurun --gpus 2)")
    L.mzb_set_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
    chk = 6 * 1048576
    n = 40 * chk + 1234567                       # 3 batches of 16 chunks, the last one ragged
    rng = np.random.default_rng(11)
    w = np.concatenate([synth.mrc_header(n, 1, 1), rng.standard_normal(n, dtype=np.float32).view(np.uint32)])
    bits = 8
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(prefix="mrcz_md_", dir=base) as d:
        d = Path(d)
        w.tofile(d / "v.mrc")
        try:
            one = (C.c_int * 1)(0)
            assert L.mzb_set_devices(one, 1) == 0
            api.zip_compress(str(d / "v.mrc"), str(d / "a.zip"), bits)
            two = (C.c_int * 2)(0, 1)
            assert L.mzb_set_devices(two, 2) == 0
            api.zip_compress(str(d / "v.mrc"), str(d / "b.zip"), bits)
            a, b = np.fromfile(d / "a.zip", np.uint8), np.fromfile(d / "b.zip", np.uint8)
            assert a.size == b.size and np.array_equal(a, b)
            api.zip_uncompress(str(d / "b.zip"), str(d / "b.out"))
            gold = oracle.erasebytes(w.view(np.uint8), bits).view(np.uint32)
            assert np.array_equal(np.fromfile(d / "b.out", np.uint32), gold)
            assert L.mzb_set_devices((C.c_int * 2)(1, 1), 2) != 0      # one worker per device
        finally:
            L.mzb_set_devices(None, 0)
