"""CPU suite (-m "not gpu"): oracle vs the reference (golden fixtures, and live when oracle/_ref is built),
the CPU model of the device codec vs zlib, the host logic, and the C-ABI exports.  No compute call
into libmrczip_b200.so here: it has no CPU path."""
import ctypes as C
import json
import zlib
from pathlib import Path

import numpy as np
import pytest

from conftest import synth_words

GOLDEN = Path(__file__).resolve().parent / "golden"


# ----------------------------------------------------------------------------- oracle: mask / split / merge
@pytest.mark.parametrize("bits", list(range(0, 33)))
def test_oracle_mask_table(oracle, bits):
    # reference workers.c:29-37
    assert oracle.mask_for_bits(bits) == ((0xFFFFFFFF << bits) & 0xFFFFFFFF if bits < 32 else 0)


def test_oracle_split_merge_roundtrip(oracle):
    w = synth_words("G", 5000)
    masked, planes = oracle.split(w, 0, True)
    assert np.array_equal(masked, w)
    assert np.array_equal(oracle.merge(planes), w)
    for j in range(4):
        assert np.array_equal(planes[j], (w >> (8 * j)).astype(np.uint8))


@pytest.mark.parametrize("bits", [1, 8, 13, 24, 32])
def test_oracle_header_exemption(oracle, bits):
    w = synth_words("R", 1000)
    masked, _ = oracle.split(w, bits, True)
    assert np.array_equal(masked[:256], w[:256])                     # workers.c:90-94
    assert np.array_equal(masked[256:], w[256:] & np.uint32(oracle.mask_for_bits(bits)))
    masked2, _ = oracle.split(w, bits, False)
    assert np.array_equal(masked2, w & np.uint32(oracle.mask_for_bits(bits)))


def test_oracle_pack_header(oracle):
    buf = (C.c_uint8 * 4)()
    for bt, ln in [(0, 0), (1, 6291456), (0, 0x7FFFFFFF), (1, 1)]:
        oracle.lib().orc_pack_header(buf, bt, ln)
        assert int.from_bytes(bytes(buf), "little") == (ln | (bt << 31))
        b2, l2 = C.c_int(), C.c_uint32()
        oracle.lib().orc_unpack_header(buf, C.byref(b2), C.byref(l2))
        assert (b2.value, l2.value) == (bt, ln)


# ----------------------------------------------------------------------------- oracle vs golden fixtures (made by the reference binary)
def _golden_cases():
    man = GOLDEN / "manifest.json"
    if not man.exists():
        return []
    return json.loads(man.read_text())["cases"]


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_oracle_matches_reference_golden(oracle, case):
    src = np.fromfile(GOLDEN / case["input"], dtype=np.uint8)
    ref_zip = np.fromfile(GOLDEN / case["ref_container"], dtype=np.uint8)
    ref_erase = np.fromfile(GOLDEN / case["ref_erasebytes"], dtype=np.uint8)
    bits = case["bits"]
    if src.size < 1024:
        # erasebytes.c:111-112 always fwrite()s 1024 bytes of its (uninitialised) buffer; only the first fsz are defined
        ref_erase = ref_erase[: src.size]
    assert np.array_equal(oracle.erasebytes(src, bits), ref_erase)          # erasebytes.c:109-134
    assert np.array_equal(oracle.compress(src, bits), ref_zip)              # byte-identical container
    assert np.array_equal(oracle.decompress(ref_zip), ref_erase[: (src.size // 4) * 4])
    assert np.array_equal(oracle.decompress(ref_zip, use_zlib=False), ref_erase[: (src.size // 4) * 4])
    for j in range(4):
        assert np.array_equal(oracle.split_file(src[: src.size // 4 * 4].view(np.uint32), bits)[1][j],
                              np.fromfile(GOLDEN / case["ref_planes"][j], dtype=np.uint8))


def test_oracle_live_against_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    assert oracle.ref_zlib_version() == "1.2.8"
    for kind, bits, n in [("G", 0, 70000), ("P", 0, 70000), ("S", 12, 30001), ("Z", 5, 9000), ("R", 31, 5000)]:
        src = synth_words(kind, n).view(np.uint8)
        assert np.array_equal(oracle.compress(src, bits), oracle.ref_compress(src, bits)), (kind, bits)
        assert np.array_equal(oracle.erasebytes(src, bits), oracle.ref_erasebytes(src, bits))
        masked, planes = oracle.ref_split(src, bits)
        m2, p2 = oracle.split_file(src.view(np.uint32), bits)
        assert np.array_equal(masked, m2) and all(np.array_equal(a, b) for a, b in zip(planes, p2))
        assert np.array_equal(oracle.ref_merge(planes), masked)
        assert np.array_equal(oracle.ref_decompress(oracle.compress(src, bits)), oracle.erasebytes(src, bits))


def test_oracle_small_chunks_and_ragged(oracle):
    # the decoder honours any chk in the file header (workers.c:578,584); ragged tails are dropped (workers.c:744)
    for nbytes in [0, 3, 4, 1024, 1027, 4096 + 2, 100001]:
        src = np.random.default_rng(nbytes).integers(0, 256, nbytes, dtype=np.uint8)
        for chk in [1000, 4096, 65536]:
            c = oracle.compress(src, 7, chk=chk)
            if nbytes < 4:
                assert c.size == 0                                      # workers.c:757-764: nothing is written
                continue
            fsz, chk2, streams = oracle.parse_container(c)
            assert (fsz, chk2) == (nbytes, chk) and len(streams) == 4 * -(-(nbytes // 4) // chk)
            assert np.array_equal(oracle.decompress(c), oracle.erasebytes(src, 7)[: nbytes // 4 * 4])


def test_oracle_inflate_is_independent_of_zlib(oracle):
    rng = np.random.default_rng(5)
    a = rng.choice([1, 2, 3, 200], 50000, p=[.5, .3, .15, .05]).astype(np.uint8)
    for strat in (zlib.Z_RLE, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY):
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, strat)
        z = co.compress(a.tobytes()) + co.flush(zlib.Z_FULL_FLUSH)
        out, used = oracle.inflate_raw(np.frombuffer(z, np.uint8), a.size)
        assert np.array_equal(out, a) and used == len(z)


# ----------------------------------------------------------------------------- CPU model of the device codec
def _plane_cases():
    rng = np.random.default_rng(0)
    fib = [1, 1]
    while sum(fib) < 16000:
        fib.append(fib[-1] + fib[-2])
    cases = {
        "zeros": np.zeros(50000, np.uint8),
        "one": np.array([7], np.uint8),
        "two": np.array([7, 7], np.uint8),
        "three_same": np.array([9, 9, 9, 9], np.uint8),
        "short": np.arange(37, dtype=np.uint8),
        "rand": rng.integers(0, 256, 40000).astype(np.uint8),
        "skew": rng.choice([0x3f, 0xbf, 0x3e, 0xbe, 0x40, 0xc0, 0x3d], 40000, p=[.3, .3, .15, .15, .04, .04, .02]).astype(np.uint8),
        "runs": np.repeat(rng.integers(0, 256, 3000).astype(np.uint8), rng.integers(1, 600, 3000))[:100000],
        "geo": np.minimum(rng.geometric(0.5, 60000), 255).astype(np.uint8),
        "fib": np.concatenate([np.full(f, i, np.uint8) for i, f in enumerate(fib)])[rng.permutation(sum(fib))][:16384],
        "exact16k": rng.choice([1, 2, 3], 16384).astype(np.uint8),
        "16k+1": rng.choice([1, 2, 3], 16385).astype(np.uint8),
        "run258": np.concatenate([np.full(259, 5, np.uint8), np.full(260, 6, np.uint8), np.full(517, 7, np.uint8), [1, 2]]).astype(np.uint8),
        "mid_entropy": rng.integers(0, 200, 33000).astype(np.uint8),
        # halving probabilities: an unlimited Huffman code would reach 20+ bits; the encoder stops at FZ_MAX_CODE_BITS (12)
        "skewed_long_codes": rng.choice(np.arange(30), 70000, p=np.r_[0.5 ** np.arange(1, 30), 0.5 ** 29]).astype(np.uint8),
    }
    g = synth_words("G", 65536).view(np.uint8).reshape(-1, 4)
    p = synth_words("P", 65536).view(np.uint8).reshape(-1, 4)
    for j in range(4):
        cases[f"G_p{j}"] = g[:, j].copy()
        cases[f"P_p{j}"] = p[:, j].copy()
    return cases


@pytest.mark.parametrize("name", sorted(_plane_cases().keys()))
def test_model_encoder_matches_sequential_rule_and_zlib(hostmodel, oracle, name):
    """The encoder (fz_enc2.cuh, what the kernels run): the device source under the 32-thread warp model produces the
    very bytes of a plain sequential restatement of the token rule, and those inflate with zlib."""
    a = _plane_cases()[name]
    c, ns = hostmodel.encode_stream_v2(a, sequential=False, skip=True)
    cs, ns_s = hostmodel.encode_stream_v2(a, sequential=True)
    assert ns == ns_s and np.array_equal(c, cs)
    d = zlib.decompressobj(-15)
    assert d.decompress(c.tobytes()) == a.tobytes()
    assert not d.eof and d.unused_data == b""
    out, used = oracle.inflate_raw(c, a.size)
    assert np.array_equal(out, a) and used == c.size
    assert c[-4:].tobytes() == b"\x00\x00\xff\xff"
    rc, o, used = hostmodel.inflate(c, a.size)
    assert rc == 0 and np.array_equal(o, a) and used == c.size
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_RLE)
    z = co.compress(a.tobytes()) + co.flush(zlib.Z_FULL_FLUSH)
    nsub = a.size // hostmodel.SUB + 1
    assert c.size <= 1.09 * len(z) + 128 * nsub or c.size <= len(z) + 0.05 * a.size


@pytest.mark.parametrize("name", ["G_p3", "P_p2", "mid_entropy", "skewed_long_codes", "run258", "16k+1"])
def test_model_encoder_with_sampled_histogram_and_standins(hostmodel, oracle, name):
    """What the kernels do: the group's code comes from every fourth sub-block, and every symbol nobody counted gets a
    stand-in count of 1 -- outside the sort and the tree, the longest codes (13..15 bits) while the symbols that were
    seen stay within 12 bits (fz_ph_standins).  Symbols that occur only in unsampled sub-blocks must still be coded."""
    a = _plane_cases()[name].copy()
    if a.size > 3 * hostmodel.SUB:
        a[hostmodel.SUB + 5: hostmodel.SUB + 40] = np.arange(200, 235, dtype=np.uint8)   # bytes only an unsampled sub-block holds
    hostmodel.set_hist_sample(4)
    try:
        c, ns = hostmodel.encode_stream_v2(a, sequential=False, skip=True)
        cs, ns_s = hostmodel.encode_stream_v2(a, sequential=True)
    finally:
        hostmodel.set_hist_sample(0)
    assert ns == ns_s and np.array_equal(c, cs)
    d = zlib.decompressobj(-15)
    assert d.decompress(c.tobytes()) == a.tobytes()       # zlib accepts the code: complete, nothing over-subscribed
    assert not d.eof and d.unused_data == b""
    rc, o, used = hostmodel.inflate(c, a.size)
    assert rc == 0 and np.array_equal(o, a) and used == c.size
    c0, _ = hostmodel.encode_stream_v2(a, sequential=False, skip=True)      # the code built from every sub-block
    nsub = a.size // hostmodel.SUB + 1
    # close to the exact code -- or, where a quarter of a nearly flat histogram is too noisy to gain 3 %, stored blocks
    assert c.size <= 1.03 * c0.size + 64 * nsub or c.size <= a.size + 15 * nsub


def test_model_nearly_incompressible_plane_costs_at_most_the_gain_threshold(hostmodel):
    """A sub-block (and a group) is coded only if that saves at least 1 / 2^FZ_MIN_GAIN_SHIFT = 3 % (nearly incompressible
    bytes decode at one symbol per table hit for a gain nobody would miss), and a sample entropy above 7.85 bits stores it
    unseen: planes zlib would shrink by 1..3 % -- typical mid-mantissa bytes -- come out stored.  The price is bounded: never
    more than 3 % (+ framing) above zlib's size, and never above the stored size."""
    rng = np.random.default_rng(3)
    n = 6 * hostmodel.SUB
    for bits_of_entropy, p_extra in ((7.8, 0.18), (7.6, 0.30), (7.95, 0.05)):
        # a flat byte distribution with a bump: p_extra of the mass on 16 symbols
        a = rng.integers(0, 256, n).astype(np.uint8)
        bump = rng.random(n) < p_extra
        a[bump] = rng.integers(0, 16, int(bump.sum())).astype(np.uint8)
        hostmodel.set_hist_sample(4)
        try:
            c, ns = hostmodel.encode_stream(a)
        finally:
            hostmodel.set_hist_sample(0)
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_RLE)
        z = co.compress(a.tobytes()) + co.flush(zlib.Z_FULL_FLUSH)
        nsub = n // hostmodel.SUB
        assert zlib.decompressobj(-15).decompress(c.tobytes()) == a.tobytes()
        assert c.size <= n + 15 * nsub + 16, (bits_of_entropy, c.size, n)                     # never above stored
        assert c.size <= len(z) * 1.035 + 80 * nsub, (bits_of_entropy, c.size, len(z))        # at most the threshold above zlib


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 511, 512, 513, 1000, 8191, 16383])
def test_model_encoder_ragged_sizes(hostmodel, n):
    """ragged sub-blocks (the last one of a file, odd chunk sizes): runs that end exactly at, before and after lane
    and step boundaries"""
    rng = np.random.default_rng(n)
    for a in (np.zeros(n, np.uint8), rng.choice([3, 4], n, p=[.9, .1]).astype(np.uint8),
              np.repeat(rng.integers(0, 256, n // 9 + 1).astype(np.uint8), 9)[:n]):
        c, _ = hostmodel.encode_stream_v2(a, sequential=False, skip=False)
        cs, _ = hostmodel.encode_stream_v2(a, sequential=True)
        assert np.array_equal(c, cs)
        assert zlib.decompressobj(-15).decompress(c.tobytes()) == a.tobytes()


def test_model_encoder_runs_across_every_boundary(hostmodel):
    """long runs of 258 k + r bytes placed so that their ends and 258-byte cuts fall on every lane position"""
    rng = np.random.default_rng(3)
    parts = []
    for i in range(40):
        parts.append(rng.integers(0, 256, int(rng.integers(1, 40))).astype(np.uint8))
        parts.append(np.full(int(rng.integers(1, 900)), int(rng.integers(0, 256)), np.uint8))
    a = np.concatenate(parts)[:3 * hostmodel.SUB]
    c, _ = hostmodel.encode_stream_v2(a, sequential=False, skip=True)
    cs, _ = hostmodel.encode_stream_v2(a, sequential=True)
    assert np.array_equal(c, cs)
    assert zlib.decompressobj(-15).decompress(c.tobytes()) == a.tobytes()


@pytest.mark.parametrize("name", sorted(_plane_cases().keys()))
def test_model_inflater_on_zlib_streams(hostmodel, name):
    a = _plane_cases()[name]
    for strat, lvl in [(zlib.Z_RLE, 6), (zlib.Z_DEFAULT_STRATEGY, 6), (zlib.Z_FIXED, 6), (zlib.Z_DEFAULT_STRATEGY, 0),
                       (zlib.Z_HUFFMAN_ONLY, 1), (zlib.Z_DEFAULT_STRATEGY, 9)]:
        co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 9, strat)
        z = co.compress(a.tobytes()) + co.flush(zlib.Z_FULL_FLUSH)
        rc, o, used = hostmodel.inflate(np.frombuffer(z, np.uint8), a.size)
        assert rc == 0 and np.array_equal(o, a) and used == len(z), (strat, lvl)


def test_model_whole_file_ratio(hostmodel, oracle):
    """Container size of the GPU encoder's format (CPU model) next to the reference's, same chunking: within 5 %."""
    from datacompressionfloat_b200 import synth
    for w, bits in [(synth_words("G", 1 << 19), 0), (synth_words("G", 1 << 19), 8), (synth_words("G", 1 << 19), 16),
                    (synth_words("P", 1 << 19), 0), (synth.mrc_volume("S", (32, 128, 128)), 12)]:
        _, planes = oracle.split_file(w, bits)
        ours = 0
        for p in planes:
            c, _ = hostmodel.encode_stream(p)
            ours += (c.size if p.size > c.size + 4 else p.size) + 4     # RAW rule, zip.c:177
        ref = oracle.compress(w.view(np.uint8), bits).size - 17
        assert ours <= 1.05 * ref, (bits, ours, ref)


def test_model_inflater_rejects_garbage(hostmodel):
    rng = np.random.default_rng(3)
    a = rng.choice([1, 2, 3], 5000).astype(np.uint8)
    c, _ = hostmodel.encode_stream(a)
    rc, o, used = hostmodel.inflate(c[: c.size // 2], a.size)      # truncated
    assert rc != 0 or o.size != a.size
    bad = c.copy()
    bad[0] |= 0x06                                                 # block type 3
    rc, _, _ = hostmodel.inflate(bad, a.size)
    assert rc != 0
    rc, o, _ = hostmodel.inflate(c, a.size - 10)                   # output too small
    assert rc != 0


def test_model_no_false_sync_markers(hostmodel):
    """The inflater finds sub-blocks by the marker 00 00 FF FF, so the encoder must never show it elsewhere."""
    MARK = bytes([0, 0, 0xFF, 0xFF])
    rng = np.random.default_rng(11)
    # (a) the check the emit stage runs on a coded fragment
    frag = rng.integers(1, 255, 5000).astype(np.uint8)
    frag[-4:] = np.frombuffer(MARK, np.uint8)
    assert not hostmodel.check_marker(frag)
    for pos in (0, 1, 2, 3, 777, 4990, 4995):
        bad = frag.copy()
        bad[pos:pos + 4] = np.frombuffer(MARK, np.uint8)
        assert hostmodel.check_marker(bad), pos
    # (b) the stored form: two stored blocks whose split breaks a marker inside the raw bytes
    for n in (1, 2, 3, 5, 254, 255, 256, 257, 509, 510, 511, 512, 8192, 16384):
        for pos in (None, 0, 1, 2, 3, 120, 251, 252, 253, 254, 255, 256, max(0, n // 2 - 2), max(0, n - 4)):
            a = rng.integers(1, 255, n).astype(np.uint8)
            a[0] = 0xFF
            if pos is not None and pos + 4 <= n:
                a[pos:pos + 4] = np.frombuffer(MARK, np.uint8)
            st = hostmodel.put_stored(a)
            assert st.size == (n + 15 if n >= 2 else n + 10)
            assert zlib.decompressobj(-15).decompress(st.tobytes()) == a.tobytes()
            assert st.tobytes().find(MARK) == st.size - 4, (n, pos)
    # (c) whole streams: markers only at sub-block ends
    for name in ("rand", "zeros", "skew", "runs", "G_p3", "mid_entropy"):
        a = _plane_cases()[name]
        c, _ = hostmodel.encode_stream(a)
        nsub = -(-a.size // hostmodel.SUB)
        assert c.tobytes().count(MARK) == nsub, name


def test_model_bfinal_and_history_rules(hostmodel):
    a = np.tile(np.arange(50, dtype=np.uint8), 400)                # long-distance matches under the default strategy
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 9)
    z = co.compress(a.tobytes()) + co.flush(zlib.Z_FINISH)         # BFINAL = 1
    rc, o, used = hostmodel.inflate(np.frombuffer(z, np.uint8), a.size)
    assert rc == 0 and np.array_equal(o, a) and used == len(z)


# ----------------------------------------------------------------------------- host logic + C ABI
def test_abi_exports_every_declared_symbol():
    from datacompressionfloat_b200 import lib
    L = lib.load()
    for sym in lib.EXPORTS:
        assert hasattr(L, sym), sym
    assert b"sm_100a" in L.mzb_version()
    assert lib.strerror(lib.E_FORMAT).startswith("malformed")
    # declared in the header <-> listed in EXPORTS
    import re
    hdr = (Path(__file__).resolve().parent.parent / "include" / "mrczip_b200.h").read_text()
    declared = set(re.findall(r"\b(mzb_\w+|run_\w+|zip_\w+|pack_header|unpack_header|\w+_context\w*|\w+_mrczip_header|get_file_size|now_sec)\s*\(", hdr))
    declared -= {"mzb_ctx"}
    assert declared <= set(lib.EXPORTS), declared - set(lib.EXPORTS)


def test_abi_pack_header_matches_oracle(oracle):
    from datacompressionfloat_b200 import lib
    L = lib.load()
    buf = C.create_string_buffer(4)
    for bt, ln in [(0, 0), (1, 6291456), (0, 0x7FFFFFFF), (1, 77)]:
        L.pack_header(buf, bt, ln)
        assert int.from_bytes(buf.raw, "little") == (ln | (bt << 31))
        b2, l2 = C.c_int(), C.c_uint32()
        L.unpack_header(buf, C.byref(b2), C.byref(l2))
        assert (b2.value, l2.value) == (bt, ln)


def test_compress_bound_and_header_helpers(oracle):
    from datacompressionfloat_b200 import Codec, file_header, chunk_range, segment_offsets, CHUNK_WORDS
    assert Codec.compress_bound(10, 4) >= 17 + 3 * 16 + 40
    h = file_header(67109888)
    assert h.size == 17 and int(h[:8].view(np.uint64)[0]) == 67109888 and int(h[8:12].view(np.uint32)[0]) == CHUNK_WORDS
    # the header the oracle (== reference) writes for the same file
    c = oracle.compress(np.zeros(4096, np.uint8), 0)
    assert np.array_equal(c[:17], file_header(4096))
    assert [chunk_range(171, r, 8) for r in range(8)][-1] == (154, 171)
    assert sum(hi - lo for lo, hi in (chunk_range(5, r, 8) for r in range(8))) == 5
    assert segment_offsets([10, 20, 5]) == ([17, 27, 47], 52)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from datacompressionfloat_b200 import Codec, MzbError
    with pytest.raises(MzbError):
        Codec(0)


# ----------------------------------------------------------------------------- block-parallel inflate of zlib-made streams (CPU model)
def test_model_blockpar_inflate_of_reference_streams(hostmodel):
    """Candidate search + measure + chain + write (fz_blockpar.cuh) on streams made with the reference's zlib
    parameters; streams with other distances must be refused (the GPU then takes the serial inflater)."""
    rng = np.random.default_rng(21)

    def z(a, strat):
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, strat)
        return np.frombuffer(co.compress(a.tobytes()) + co.flush(zlib.Z_FULL_FLUSH), np.uint8)

    cases = {
        "exp": synth_words("G", 150000).view(np.uint8).reshape(-1, 4)[:, 3].copy(),       # several blocks
        "counts": synth_words("P", 150000).view(np.uint8).reshape(-1, 4)[:, 2].copy(),
        "zeros": np.zeros(300000, np.uint8),                                                 # run continues across blocks
        "runs": np.repeat(rng.integers(0, 256, 2000).astype(np.uint8), rng.integers(1, 400, 2000))[:200000],
        "stored_mix": np.concatenate([rng.integers(0, 256, 70000).astype(np.uint8), rng.choice([1, 2, 3], 90000).astype(np.uint8),
                                      rng.integers(0, 256, 40000).astype(np.uint8)]),
    }
    for name, a in cases.items():
        c = z(a, zlib.Z_RLE)
        rc, o, ncand, noff = hostmodel.inflate_blockpar(c, a.size)
        assert rc == 0 and np.array_equal(o, a), (name, rc)
        assert ncand >= 1
    # fixed-Huffman block (zlib picks it for tiny inputs): no header to search for, measured by the chain walk itself
    tiny = np.array([5, 5, 5, 5, 5, 9], np.uint8)
    rc, o, ncand, _ = hostmodel.inflate_blockpar(z(tiny, zlib.Z_RLE), tiny.size)
    assert rc == 0 and ncand == 0 and np.array_equal(o, tiny)
    # a finished stream (Z_FINISH): its last block has BFINAL = 1, which the header search skips -- the chain walk
    # measures that block itself
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_RLE)
    c = np.frombuffer(co.compress(cases["exp"].tobytes()) + co.flush(zlib.Z_FINISH), np.uint8)
    rc, o, _, _ = hostmodel.inflate_blockpar(c, cases["exp"].size)
    assert rc == 0 and np.array_equal(o, cases["exp"])
    # default strategy = long distances: refused, never wrong
    a = np.tile(np.arange(200, dtype=np.uint8), 600)
    rc, _, _, _ = hostmodel.inflate_blockpar(z(a, zlib.Z_DEFAULT_STRATEGY), a.size)
    assert rc < 0
    # truncated stream: refused
    c = z(cases["exp"], zlib.Z_RLE)
    rc, _, _, _ = hostmodel.inflate_blockpar(c[: c.size // 2], cases["exp"].size)
    assert rc != 0


def test_model_warp_built_lookup_table_equals_the_searched_one(hostmodel):
    """The block-parallel decoder's table is built by interval walk + in-table packing (fz_lut_fill_lane / fz_lut_pack)
    instead of one canonical search per entry: entry for entry the same table, for complete, incomplete, skewed, flat and
    single-symbol codes."""
    rng = np.random.default_rng(11)

    def lengths_from_freq(f):          # a Huffman code for frequencies f through zlib-like limiting: use package-free heuristic
        import heapq
        n = len(f)
        heap = [(int(w), i, None, None) for i, w in enumerate(f) if w > 0]
        lens = np.zeros(n, np.int64)
        if len(heap) == 1:
            lens[heap[0][1]] = 1
            return lens
        heapq.heapify(heap)
        nodes = {}
        uid = n
        while len(heap) > 1:
            a = heapq.heappop(heap); b = heapq.heappop(heap)
            nodes[uid] = (a, b)
            heapq.heappush(heap, (a[0] + b[0], uid, a, b))
            uid += 1
        stack = [(heap[0], 0)]
        while stack:
            (w, i, a, b), d = stack.pop()
            if a is None:
                lens[i] = max(d, 1)
            else:
                stack.append((a, d + 1)); stack.append((b, d + 1))
        return lens

    cases = []
    for trial in range(40):
        nsym = int(rng.integers(2, 287))
        syms = rng.choice(286, nsym, replace=False)
        f = np.zeros(288, np.int64)
        kind = trial % 4
        if kind == 0:
            f[syms] = rng.integers(1, 1000, nsym)
        elif kind == 1:
            f[syms] = (2.0 ** rng.integers(0, 14, nsym)).astype(np.int64)
        elif kind == 2:
            f[syms] = 1
        else:
            f[syms] = np.maximum(1, (1e6 * 0.6 ** np.arange(nsym))).astype(np.int64)
        f[256] = max(f[256], 1)
        lens = lengths_from_freq(f)
        if lens.max() > 15:
            continue
        cases.append(lens.astype(np.uint8))
    one = np.zeros(288, np.uint8); one[65] = 1                       # a single code of one bit: incomplete
    two = np.zeros(288, np.uint8); two[0] = 1; two[256] = 1
    fixed = np.r_[np.full(144, 8), np.full(112, 9), np.full(24, 7), np.full(8, 8)].astype(np.uint8)   # the fixed code of RFC 1951
    gap = np.zeros(288, np.uint8); gap[[1, 2, 3]] = 2; gap[256] = 5   # incomplete: patterns without a code
    cases += [one, two, fixed, gap]
    assert len(cases) > 20
    for lens in cases:
        assert hostmodel.lut_compare(lens) == 0, lens.tolist()


def test_model_blockpar_tile_records_and_pool_exhaustion(hostmodel):
    """The measure pass leaves tile records so that the write pass decodes every sub-range once; when the pool is
    too small the write pass searches again -- same bytes either way."""
    a = synth_words("G", 400000).view(np.uint8).reshape(-1, 4)[:, 3].copy()
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_RLE)
    c = np.frombuffer(co.compress(a.tobytes()) + co.flush(zlib.Z_FULL_FLUSH), np.uint8)
    try:
        for cap, expect_all in [(4096, True), (0, False), (5, False)]:
            hostmodel.set_tile_pool_cap(cap)
            rc, o, ncand, _ = hostmodel.inflate_blockpar(c, a.size)
            assert rc == 0 and np.array_equal(o, a), cap
            assert (hostmodel.table_blocks() == ncand) == expect_all, (cap, hostmodel.table_blocks(), ncand)
    finally:
        hostmodel.set_tile_pool_cap(4096)


def test_mrc_header_parse_needs_no_gpu():
    """mzb_mrc_parse (reference src/tool/mrcviewer.c:20-71: nx, ny, nz, mod at words 0..3, next at word 23)"""
    import ctypes as C
    from datacompressionfloat_b200 import lib
    L = lib.load()

    class Info(C.Structure):
        _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("mode", C.c_int32), ("next", C.c_int32),
                    ("is_float32", C.c_int32), ("data_offset", C.c_uint64)]
    h = np.zeros(256, dtype=np.int32)
    h[0:3] = (4096, 4096, 40)
    h[3] = 2
    h[23] = 131072
    info = Info()
    assert L.mzb_mrc_parse(C.c_void_p(h.ctypes.data), C.c_size_t(1024), C.byref(info)) == 0
    assert (info.nx, info.ny, info.nz, info.mode, info.next, info.is_float32, info.data_offset) == (4096, 4096, 40, 2, 131072, 1, 1024 + 131072)
    h[3] = 6
    assert L.mzb_mrc_parse(C.c_void_p(h.ctypes.data), C.c_size_t(1024), C.byref(info)) == 0 and info.is_float32 == 0
    h[3] = 77
    assert L.mzb_mrc_parse(C.c_void_p(h.ctypes.data), C.c_size_t(1024), C.byref(info)) == lib.E_FORMAT
    h[3] = 2; h[1] = 0
    assert L.mzb_mrc_parse(C.c_void_p(h.ctypes.data), C.c_size_t(1024), C.byref(info)) == lib.E_FORMAT
    assert L.mzb_mrc_parse(C.c_void_p(h.ctypes.data), C.c_size_t(512), C.byref(info)) == lib.E_ARG
