"""Developer probe (GPU box): per-stage device times of the hot path on synthetic volumes.
Usage: python tools/perf_probe.py [--gib 1] [--cases G0,G8,G16,P0]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from datacompressionfloat_b200 import Codec, CHUNK_WORDS  # noqa: E402


def gen(kind, nwords, device="cuda"):
    g = torch.Generator(device=device)
    g.manual_seed({"G": 1234, "P": 4321, "S": 7}[kind])
    if kind == "G":
        d = torch.randn(nwords, generator=g, device=device, dtype=torch.float32)
    elif kind == "P":
        d = torch.poisson(torch.full((nwords,), 2.0, device=device), generator=g)
    else:
        x = torch.linspace(0, 4000 * np.pi, nwords, device=device)
        d = torch.sin(x) * torch.cos(x / 7) + 0.25 * torch.randn(nwords, generator=g, device=device)
    w = d.view(torch.int32)
    w[:256] = 0
    w[0] = nwords - 256
    w[3] = 2
    return w


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--cases", default="G0,G8,G16,P0,S12")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--no-split", action="store_true")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    nwords = int(a.gib * (1 << 30)) // 4
    codec = Codec.on_current_stream(batch_chunks=a.batch)
    codec.set_profiling(True)
    out = []
    # ---- split / merge alone (config 2): 8 algorithmic bytes per word
    w = gen("G", nwords)
    planes = torch.empty((4, (nwords + 255) // 256 * 256), dtype=torch.uint8, device="cuda")
    back = torch.empty(nwords, dtype=torch.int32, device="cuda")
    for v in (() if a.no_split else (0, 1)):
        codec.set_variant(v, v)
        t_s, _ = timeit(lambda: codec.mask_split(w, 8, 256, out=planes))
        t_m, _ = timeit(lambda: codec.merge(planes, nwords, out=back))
        rec = dict(kind="split_merge", variant=v, gib=a.gib, split_ms=t_s, merge_ms=t_m,
                   split_GBs=nwords * 8 / t_s / 1e6, merge_GBs=nwords * 8 / t_m / 1e6)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    codec.set_variant(0, 0)
    # torch copy as the local HBM yardstick (same definition as MEASURED_PEAKS.json: read + write bytes)
    src = w
    dst = torch.empty_like(w)
    t_c, _ = timeit(lambda: dst.copy_(src))
    print(json.dumps(dict(kind="torch_copy", GBs=nwords * 8 / t_c / 1e6, ms=t_c)), flush=True)
    del planes, back, dst
    # ---- full pipeline
    for case in a.cases.split(","):
        kind, bits = case[0], int(case[1:])
        w = gen(kind, nwords)
        cont_buf = torch.empty(Codec.compress_bound(nwords), dtype=torch.uint8, device="cuda")
        outw = torch.empty(nwords, dtype=torch.int32, device="cuda")
        holder = {}

        def comp():
            holder["c"] = codec.compress(w, bits, out=cont_buf)

        def decomp():
            holder["d"] = codec.decompress(holder["c"], out=outw)

        tc, tc_med = timeit(comp, iters=a.iters, warm=1)
        st_c, ms_c = codec.stats(), codec.stage_ms()
        td, td_med = timeit(decomp, iters=a.iters, warm=1)
        st_d, ms_d = codec.stats(), codec.stage_ms()
        mask = -1 << bits if bits < 32 else 0
        ref = w.clone()
        ref[256:] &= mask
        ok = bool(torch.equal(ref, holder["d"]))
        nb = nwords * 4
        rec = dict(kind="pipeline", case=case, gib=a.gib, ok=ok, ratio=(holder["c"].numel() - 17) / nb,
                   comp_ms=tc, decomp_ms=td, comp_GBs=nb / tc / 1e6, decomp_GBs=nb / td / 1e6,
                   roundtrip_GBs=nb / (tc + td) / 1e6,
                   comp_stage_ms={k: round(v, 3) for k, v in ms_c.items() if v > 0},
                   decomp_stage_ms={k: round(v, 3) for k, v in ms_d.items() if v > 0},
                   raw_streams=st_c["raw_streams"], stored_sub=st_c["stored_subblocks"], streams=st_c["streams"],
                   general=st_d["general_streams"], fast_failed=st_d["fast_failed"])
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del w, cont_buf, outw, ref, holder
        torch.cuda.empty_cache()
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/perf_probe.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
