"""Runs the BASELINE.json configs that fit one GPU box and prints one JSON line per config
(results are copied into profiles/README.md).  GPU box only.

  config 1  256^3 MRC volume (64 MiB), G/P/S x b: GPU file-to-file (zip_compress / zip_uncompress of the C ABI)
            next to the reference's mrc_tar_c on the host (single thread), outputs cross-checked both ways
  config 2  1 GiB array: mask + split only / merge only, planes bit-exact vs the oracle on a slice
  config 4  16 GiB stream, device resident (4 batches of 192 chunks), vs the reference pthread pool sample
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from datacompressionfloat_b200 import Codec, synth, zip_compress, zip_uncompress  # noqa: E402
from oracle import oracle as O  # noqa: E402


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def config1():
    out = []
    tmp = tempfile.mkdtemp(prefix="cfg1_", dir="/dev/shm")
    for kind, bits in [("G", 0), ("G", 8), ("G", 16), ("P", 0), ("S", 0), ("S", 12)]:
        w = synth.mrc_volume(kind, (256, 256, 256))
        src, z, back = f"{tmp}/v.mrc", f"{tmp}/v.mrc.zip", f"{tmp}/v.out"
        w.tofile(src)
        golden = O.erasebytes(w.view(np.uint8), bits)
        zip_compress(src, z, bits); zip_uncompress(z, back)          # warm (context creation, pinned buffers)
        t0 = time.perf_counter(); zip_compress(src, z, bits); t_c = time.perf_counter() - t0
        t0 = time.perf_counter(); zip_uncompress(z, back); t_d = time.perf_counter() - t0
        ok_rt = np.array_equal(np.fromfile(back, dtype=np.uint8), golden)
        gsize = os.path.getsize(z)
        rec = dict(config=1, kind=kind, bits=bits, bytes=w.size * 4, gpu_zip_s=t_c, gpu_unzip_s=t_d,
                   gpu_file_to_file_GBs=w.size * 4 / (t_c + t_d) / 1e9, gpu_ratio=(gsize - 17) / (w.size * 4), gpu_roundtrip_ok=ok_rt)
        if O.have_ref():
            rz, rback = f"{tmp}/r.zip", f"{tmp}/r.out"
            t0 = time.perf_counter()
            subprocess.run([O.REF_DIR / "mrc_tar_c", "-i", src, "-o", rz, "-b", str(bits), "-t", "zip"], check=True, stdout=subprocess.DEVNULL)
            t_rc = time.perf_counter() - t0
            t0 = time.perf_counter()
            subprocess.run([O.REF_DIR / "mrc_tar_c", "-i", rz, "-o", rback, "-t", "unzip"], check=True, stdout=subprocess.DEVNULL)
            t_rd = time.perf_counter() - t0
            # cross decode: reference binary reads the GPU file; GPU reads the reference file
            subprocess.run([O.REF_DIR / "mrc_tar_c", "-i", z, "-o", rback, "-t", "unzip"], check=True, stdout=subprocess.DEVNULL)
            ok_ref_reads_gpu = np.array_equal(np.fromfile(rback, dtype=np.uint8), golden)
            zip_uncompress(rz, back)
            ok_gpu_reads_ref = np.array_equal(np.fromfile(back, dtype=np.uint8), golden)
            rec.update(ref_zip_s=t_rc, ref_unzip_s=t_rd, ref_1thread_GBs=w.size * 4 / (t_rc + t_rd) / 1e9,
                       ref_ratio=(os.path.getsize(rz) - 17) / (w.size * 4), ref_reads_gpu_ok=ok_ref_reads_gpu, gpu_reads_ref_ok=ok_gpu_reads_ref)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    subprocess.run(["rm", "-rf", tmp])
    return out


def config2():
    n = (1 << 30) // 4
    codec = Codec.on_current_stream()
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    w = torch.randn(n, generator=g, device="cuda").view(torch.int32)
    w[:256] = 0
    planes = torch.empty((4, n), dtype=torch.uint8, device="cuda")
    back = torch.empty(n, dtype=torch.int32, device="cuda")
    recs = []
    for bits in (0, 8, 16):
        t_s = ev_time(lambda: codec.mask_split(w, bits, 256, out=planes))
        t_m = ev_time(lambda: codec.merge(planes, n, out=back))
        # bit-exact against the oracle on three slices (start incl. header exemption, middle, end)
        ok = True
        for lo in (0, n // 2 - 5000, n - 100000):
            hi = lo + 100000
            ws = w[lo:hi].cpu().numpy().view(np.uint32)
            _, pl = O.split(ws, bits, lo == 0)
            got = planes[:, lo:hi].cpu().numpy()
            ok &= all(np.array_equal(got[j], pl[j]) for j in range(4))
        ref = w.clone(); ref[256:] &= (-1 << bits) if bits < 32 else 0
        ok &= bool(torch.equal(ref, back))
        rec = dict(config=2, bits=bits, words=n, split_ms=t_s, merge_ms=t_m, split_GBs_traffic=n * 8 / t_s / 1e6,
                   merge_GBs_traffic=n * 8 / t_m / 1e6, split_frac_of_measured_peak=n * 8 / t_s / 1e6 / 6532.5,
                   merge_frac_of_measured_peak=n * 8 / t_m / 1e6 / 6532.5, bit_exact=ok)
        print(json.dumps(rec), flush=True)
        recs.append(rec)
    return recs


def config4(gib=16):
    n = int(gib * (1 << 30)) // 4 + 256
    codec = Codec.on_current_stream()
    g = torch.Generator(device="cuda"); g.manual_seed(99)
    w = torch.empty(n, dtype=torch.int32, device="cuda")
    step = 1 << 28
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        w[lo:hi] = torch.randn(hi - lo, generator=g, device="cuda").view(torch.int32)
    w[:256] = 0
    cont = torch.empty(Codec.compress_bound(n), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    bits = 8
    h = {}
    t_c = ev_time(lambda: h.__setitem__("c", codec.compress(w, bits, out=cont)), iters=3, warm=1)
    t_d = ev_time(lambda: h.__setitem__("d", codec.decompress(h["c"], out=out)), iters=3, warm=1)
    w[256:] &= (-1 << bits)
    ok = bool(torch.equal(w, h["d"]))
    rec = dict(config=4, gib=gib, bits=bits, kind="G", compress_GBs=n * 4 / t_c / 1e6, decompress_GBs=n * 4 / t_d / 1e6,
               roundtrip_GBs=n * 4 / (t_c + t_d) / 1e6, ratio=(h["c"].numel() - 17) / (n * 4), bit_exact=ok,
               mem_GiB=torch.cuda.max_memory_allocated() / 2**30)
    print(json.dumps(rec), flush=True)
    return [rec]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,4")
    a = ap.parse_args()
    allr = []
    for c in a.configs.split(","):
        allr += {"1": config1, "2": config2, "4": config4}[c]()
    for r in allr:
        r["commit"] = os.environ.get("MRCZIP_COMMIT")
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/configs.json").write_text(json.dumps(allr, indent=1))
