"""Developer probe: zip_compress / zip_uncompress of a 4 GiB volume in /dev/shm for several I/O thread counts
(mzb_set_io_threads).  usage: python tools/io_sweep.py [GiB]"""
import ctypes as C, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datacompressionfloat_b200 import lib
L = lib.load()
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * 2**30 / 4)
g = torch.Generator(device="cuda"); g.manual_seed(1234)
w = torch.randn(n, generator=g, device="cuda")
d = "/dev/shm/mrczip_io"
os.makedirs(d, exist_ok=True)
src, z, out = f"{d}/v.mrc", f"{d}/v.mrc.zip", f"{d}/v.out"
w.cpu().numpy().tofile(src)
flag = C.c_int.in_dll(L, "isTestThroughput")
ctx = lib.CtxT()
res = {}
for t in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "4,8,16,2".split(","))]:
    L.mzb_set_io_threads(t)
    r = {}
    for rep in range(2):
        L.init_context(C.byref(ctx)); flag.value = 0
        t0 = time.perf_counter(); L.zip_compress(C.byref(ctx), src.encode(), z.encode(), 8); t1 = time.perf_counter()
        L.zip_uncompress(C.byref(ctx), z.encode(), out.encode()); t2 = time.perf_counter()
        flag.value = 1
        L.zip_uncompress(C.byref(ctx), z.encode(), out.encode()); t3 = time.perf_counter()
        L.zip_compress(C.byref(ctx), src.encode(), z.encode(), 8); t4 = time.perf_counter()
        flag.value = 0
        r = {"zip_GBs": round(n * 4 / (t1 - t0) / 1e9, 2), "unzip_write_GBs": round(n * 4 / (t2 - t1) / 1e9, 2),
             "unzip_d1_GBs": round(n * 4 / (t3 - t2) / 1e9, 2), "zip_d1_GBs": round(n * 4 / (t4 - t3) / 1e9, 2)}
    res[t] = r
    print(t, r, flush=True)
print(json.dumps({"io_sweep": res, "cpus": os.cpu_count(), "GiB": gib}))
for f in (src, z, out):
    if os.path.exists(f): os.remove(f)
