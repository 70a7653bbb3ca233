"""GPU drop-in suite (-m gpu): the reference's own UNMODIFIED mains (src/main/mrc_tar.c, mrc_tarx.c), linked against
libmrczip_b200.so instead of the reference's src/core objects (oracle/Makefile `dropin`), run the reference's own
test procedure (script/run_full_test.sh:84-108): for b = 0..N  unzip(zip(x, b)) == erasebytes(x, b), byte-exact --
plus BASELINE.json configs[0] at its named size (256^3, G / P / S, b in {0, 8, 16}) in both directions against the
reference binaries, and a device-API run that crosses kernel batches at the reference chunk size with EVERY
compressed payload inflated by the reference's libz.
"""
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
N256 = 256 ** 3


def _need(*names):
    for n in names:
        if not (REF / n).exists():
            pytest.skip(f"oracle/_ref/{n} not present (built here from /root/reference by oracle/Makefile)")


def _run(*args):
    subprocess.run([str(a) for a in args], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


@pytest.fixture(scope="module")
def shm():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    with tempfile.TemporaryDirectory(prefix="mrcz_dropin_", dir=base) as d:
        yield Path(d)


@pytest.fixture(scope="module")
def volumes(shm):
    """256^3 MRC volumes (1024-byte header + 16,777,216 float32: 2 chunks + one of 4,194,560 words) of the three
    distributions, and their erasebytes goldens made by the reference's own generator on demand."""
    from datacompressionfloat_b200 import synth
    out = {}
    for kind in ("G", "P", "S"):
        w = synth.mrc_volume(kind, (256, 256, 256))          # the volumes BASELINE.md section 2 names (its seeds)
        p = shm / f"{kind}.mrc"
        w.tofile(p)
        out[kind] = (p, w)
    return out


def _golden(shm, volumes, kind, bits):
    g = shm / f"{kind}_mask{bits}bits.mrc"
    if not g.exists():
        _run(REF / "erasebytes_c", "-i", volumes[kind][0], "-o", g, "-b", bits)
    return g


def _same(a: Path, b: Path) -> bool:
    return subprocess.run(["cmp", "-s", str(a), str(b)]).returncode == 0


@pytest.mark.parametrize("kind,bit_list", [("G", list(range(0, 9))), ("P", [0, 8]), ("S", [0, 8])])
def test_reference_main_linked_to_the_library_runs_the_reference_procedure(shm, volumes, kind, bit_list):
    """script/run_full_test.sh:84-108 with the reference's mrc_tar main on top of libmrczip_b200.so"""
    _need("mrc_tar_dropin", "erasebytes_c")
    src = volumes[kind][0]
    for bits in bit_list:
        z, back = shm / f"dz_{kind}_{bits}.zip", shm / f"dz_{kind}_{bits}.mrc"
        _run(REF / "mrc_tar_dropin", "-i", src, "-o", z, "-b", bits, "-t", "zip")
        _run(REF / "mrc_tar_dropin", "-i", z, "-o", back, "-t", "unzip")
        assert _same(_golden(shm, volumes, kind, bits), back), (kind, bits)
        back.unlink()
        if bits != 8:
            z.unlink()


@pytest.mark.parametrize("kind", ["G", "P", "S"])
def test_drop_in_and_reference_decode_each_other(shm, volumes, kind):
    _need("mrc_tar_dropin", "mrc_tar_c", "erasebytes_c")
    src, bits = volumes[kind][0], 8
    gold = _golden(shm, volumes, kind, bits)
    zd, zr = shm / f"x_{kind}_d.zip", shm / f"x_{kind}_r.zip"
    _run(REF / "mrc_tar_dropin", "-i", src, "-o", zd, "-b", bits, "-t", "zip")
    _run(REF / "mrc_tar_c", "-i", src, "-o", zr, "-b", bits, "-t", "zip")
    # the container header and chunk structure are the reference's; the size is within 5 % of its
    assert np.array_equal(np.fromfile(zd, np.uint8, 17), np.fromfile(zr, np.uint8, 17))
    assert zd.stat().st_size <= 1.05 * zr.stat().st_size
    b1, b2 = shm / f"x_{kind}_1.mrc", shm / f"x_{kind}_2.mrc"
    _run(REF / "mrc_tar_c", "-i", zd, "-o", b1, "-t", "unzip")          # reference code + its libz 1.2.8 on our streams
    _run(REF / "mrc_tar_dropin", "-i", zr, "-o", b2, "-t", "unzip")     # GPU on the reference's streams
    assert _same(gold, b1) and _same(gold, b2)
    for f in (zd, zr, b1, b2):
        f.unlink()


def test_reference_multi_file_main_linked_to_the_library(shm, volumes):
    """mrc_tarx (file list, N pthread workers, DIR/<name>.mrc.zip naming: mrc_tarx.c:216-256, adapt.c:266-320)"""
    _need("mrc_tarx_dropin", "erasebytes_c")
    names = []
    for i, kind in enumerate(["G", "P", "S", "G"]):
        p = shm / f"stack{i}.mrc"
        if not p.exists():
            os.link(volumes[kind][0], p)
        names.append((p, kind))
    lst, zdir, odir = shm / "zip.txt", shm / "zipped", shm / "unzipped"
    zdir.mkdir(); odir.mkdir()
    lst.write_text("".join(f"{p}\n" for p, _ in names))
    _run(REF / "mrc_tarx_dropin", "-i", lst, "-t", "zip", "-o", zdir, "-b", 8, "-n", 2)
    zips = [zdir / (p.name + ".zip") for p, _ in names]
    assert all(z.exists() for z in zips)
    ulst = shm / "unzip.txt"
    ulst.write_text("".join(f"{z}\n" for z in zips))
    _run(REF / "mrc_tarx_dropin", "-i", ulst, "-t", "unzip", "-o", odir, "-n", 2)
    for (p, kind) in names:
        assert _same(_golden(shm, volumes, kind, 8), odir / p.name), p.name


# ----------------------------------------------------------------------------- BASELINE configs[0] at its named size
@pytest.mark.parametrize("kind", ["G", "P", "S"])
@pytest.mark.parametrize("bits", [0, 8, 16])
def test_config1_256cube_both_directions_against_the_reference_binaries(codec, oracle, shm, volumes, kind, bits):
    _need("mrc_tar_c", "erasebytes_c")
    src, w = volumes[kind]
    gold = np.fromfile(_golden(shm, volumes, kind, bits), dtype=np.uint32)
    d_w = torch.from_numpy(w.view(np.int32)).cuda()
    cont = codec.compress(d_w, bits)
    back = codec.decompress(cont)
    assert np.array_equal(back.cpu().numpy().view(np.uint32), gold)                   # GPU round trip == erasebytes
    zg, bg = shm / "c1_gpu.zip", shm / "c1_gpu.mrc"
    cont.cpu().numpy().tofile(zg)
    _run(REF / "mrc_tar_c", "-i", zg, "-o", bg, "-t", "unzip")                        # reference inflates the GPU container
    assert np.array_equal(np.fromfile(bg, dtype=np.uint32), gold)
    zr = shm / "c1_ref.zip"
    _run(REF / "mrc_tar_c", "-i", src, "-o", zr, "-b", bits, "-t", "zip")
    ref_cont = np.fromfile(zr, dtype=np.uint8)
    back2 = codec.decompress(torch.from_numpy(ref_cont).cuda())                       # GPU inflates the reference container
    assert np.array_equal(back2.cpu().numpy().view(np.uint32), gold)
    assert cont.numel() <= 1.05 * ref_cont.size, (cont.numel(), ref_cont.size)        # ratio within 5 % of the reference's
    for f in (zg, bg, zr):
        f.unlink()


# ----------------------------------------------------------------------------- kernel batches at the reference chunk size
def test_device_api_across_kernel_batches_every_payload_through_reference_libz(oracle):
    """5 full reference-sized chunks + a ragged one, 2 chunks per kernel batch: the running container offset, the
    per-batch stream tables and the header exemption cross three batch boundaries; every COMPRESSED payload of the
    result is inflated by the reference's mzlib_inf (its libz 1.2.8), every RAW one compared verbatim."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    from datacompressionfloat_b200 import Codec
    chk = 6 * 1048576
    w = synth_words("S", 5 * chk + 123457 - 256, seed=9)
    bits = 12
    with Codec(0, batch_chunks=2) as c2, Codec(0) as c1:
        d_w = torch.from_numpy(w.view(np.int32)).cuda()
        cont = c2.compress(d_w, bits)
        one = c1.compress(d_w, bits)
        assert torch.equal(cont, one)                        # batching does not change a byte
        back = c2.decompress(cont)
        gold = oracle.erasebytes(w.view(np.uint8), bits).view(np.uint32)
        assert np.array_equal(back.cpu().numpy().view(np.uint32), gold)
    h = cont.cpu().numpy()
    fsz, chk2, streams = oracle.parse_container(h)
    assert (fsz, chk2, len(streams)) == (w.size * 4, chk, 24)
    _, ref_planes = oracle.split_file(w, bits)
    pos = [0, 0, 0, 0]
    ncomp = 0
    for i, s in enumerate(streams):
        j = i % 4
        want = ref_planes[j][pos[j]: pos[j] + s["n"]]
        pos[j] += s["n"]
        payload = h[s["offset"]: s["offset"] + s["len"]]
        if s["raw"]:
            assert np.array_equal(payload, want)
        else:
            ncomp += 1
            assert np.array_equal(oracle.ref_inflate(payload, s["n"]), want), (i, j)
    assert ncomp >= 6
