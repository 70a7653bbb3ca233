"""Developer stress run (GPU box): many device round trips on random sizes / mask bits / distributions, every one compared
word for word with torch's own masking of the input; the lean and the full group inflater alternate.
Usage: python tools/stress_roundtrip.py [--seconds 60] [--max-mib 1536]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from datacompressionfloat_b200 import Codec  # noqa: E402


def gen(kind, n, g):
    if kind == "G":
        d = torch.randn(n, generator=g, device="cuda")
    elif kind == "P":
        d = torch.poisson(torch.full((n,), 2.0, device="cuda"), generator=g)
    elif kind == "S":
        x = torch.linspace(0, 4000 * np.pi * n / 2**30, n, device="cuda")
        d = torch.sin(x) * torch.cos(x / 7) + 0.25 * torch.randn(n, generator=g, device="cuda")
    elif kind == "Z":   # long runs of a few values with islands of noise
        d = torch.zeros(n, device="cuda")
        k = max(1, n // 50000)
        idx = torch.randint(0, n, (k,), generator=g, device="cuda")
        d[idx] = torch.randn(k, generator=g, device="cuda")
        d = torch.cumsum(d, 0).round()
    else:               # random bits
        return torch.randint(-2**31, 2**31 - 1, (n,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
    return d.view(torch.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--max-mib", type=int, default=1536)
    a = ap.parse_args()
    rng = np.random.default_rng(20261019)
    g = torch.Generator(device="cuda")
    codec = Codec.on_current_stream()
    t0, runs, bad = time.time(), 0, []
    while time.time() - t0 < a.seconds:
        kind = "GPSZR"[int(rng.integers(0, 5))]
        bits = int(rng.choice([0, 1, 4, 7, 8, 9, 12, 15, 16, 20, 23, 24, 31, 32]))
        n = int(rng.integers(1, a.max_mib * (1 << 18)))
        if rng.random() < 0.3:
            n = int(rng.integers(1, 1 << 20))
        chk = int(rng.choice([6291456, 6291456, 1 << 20, 65536, 1000, 4096 + 16]))
        if n // chk > 3000:
            chk = 6291456
        g.manual_seed(int(rng.integers(0, 2**31)))
        w = gen(kind, n, g)
        ex = min(256, n)
        ref = w.clone()
        ref[ex:] &= (-1 << bits) if bits < 32 else 0
        variant = runs & 1
        codec.set_inflate_variant(variant)
        cont = codec.compress(w, bits, chk=chk)
        back = codec.decompress(cont)
        st = codec.stats()
        ok = bool(torch.equal(ref, back)) and st["general_streams"] == 0 and st["fast_failed"] == 0
        if not ok:
            bad.append(dict(kind=kind, bits=bits, n=n, chk=chk, variant=variant, stats={k: int(v) for k, v in st.items()}))
        runs += 1
        del w, ref, cont, back
    codec.set_inflate_variant(0)
    rec = dict(runs=runs, failures=bad, seconds=round(time.time() - t0, 1))
    print(json.dumps(rec))
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/stress_roundtrip.json").write_text(json.dumps(rec, indent=1))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
