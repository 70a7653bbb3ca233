"""File-to-file throughput of the drop-in entry points (zip_compress / zip_uncompress, reference adapt.c:28-90) on a
synthetic volume in /dev/shm, with the overlapped multi-threaded I/O and with the plain fread / fwrite loop
(MRCZIP_SERIAL_IO=1).  usage: python tools/file_throughput.py [GiB] [kind] [bits]"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datacompressionfloat_b200 import zip_compress, zip_uncompress, synth

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
kind = sys.argv[2] if len(sys.argv) > 2 else "G"
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 8
n = int(gib * 2**30 / 4)
g = torch.Generator(device="cuda"); g.manual_seed(1234)
if kind == "G":
    w = torch.randn(n, generator=g, device="cuda")
else:
    w = torch.poisson(torch.full((n,), 2.0, device="cuda"), generator=g)
d = "/dev/shm/mrczip_ft"
os.makedirs(d, exist_ok=True)
src, z, out = f"{d}/v.mrc", f"{d}/v.mrc.zip", f"{d}/v.out"
w.cpu().numpy().tofile(src)
mask = np.uint32((0xFFFFFFFF << bits) & 0xFFFFFFFF)
res = {}
for mode in ("overlapped", "serial", "overlapped"):
    if mode == "serial": os.environ["MRCZIP_SERIAL_IO"] = "1"
    else: os.environ.pop("MRCZIP_SERIAL_IO", None)
    t0 = time.perf_counter(); zip_compress(src, z, bits); t1 = time.perf_counter(); zip_uncompress(z, out); t2 = time.perf_counter()
    res[mode] = {"zip_s": round(t1 - t0, 3), "unzip_s": round(t2 - t1, 3), "zip_GBs": round(n * 4 / (t1 - t0) / 1e9, 2),
                 "unzip_GBs": round(n * 4 / (t2 - t1) / 1e9, 2), "roundtrip_GBs": round(n * 4 / (t2 - t0) / 1e9, 2)}
back = np.fromfile(out, dtype=np.uint32)
orig = np.fromfile(src, dtype=np.uint32)
orig[256:] &= mask
res["bit_exact"] = bool(np.array_equal(back, orig))
res["ratio"] = round(os.path.getsize(z) / os.path.getsize(src), 4)
res["config"] = {"GiB": gib, "kind": kind, "bits": bits, "dir": d, "cpus": os.cpu_count()}
print(json.dumps(res))
for f in (src, z, out): os.remove(f)
