"""Developer probe: does running two half-volume codecs concurrently (two host threads, two streams) beat one
full-volume codec?  (Tests whether overlapping the HBM-bound and the latency-bound kernels of different batches pays.)"""
import sys, threading, time
import torch
sys.path.insert(0, ".")
from datacompressionfloat_b200 import Codec

def gen(n, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    return torch.randn(n, generator=g, device="cuda").view(torch.int32)

def run(nparts, gib=4.0, bits=8, reps=5):
    chunk = 6 * 1048576
    nchunks = int(gib * 2**30 / 4 / chunk)
    per = nchunks // nparts
    words = [gen(per * chunk, 1234 + i) for i in range(nparts)]
    codecs = [Codec(0) for _ in range(nparts)]
    outs = [torch.empty(Codec.compress_bound(w.numel()), dtype=torch.uint8, device="cuda") for w in words]
    backs = [torch.empty_like(w) for w in words]
    def work(i):
        seg = codecs[i].compress(words[i], bits, exempt_words=0, write_file_header=False, out=outs[i])
        codecs[i].decompress(seg, has_file_header=False, nwords=words[i].numel(), out=backs[i])
    def once():
        ts = [threading.Thread(target=work, args=(i,)) for i in range(nparts)]
        for t in ts: t.start()
        for t in ts: t.join()
    for _ in range(2): once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    tot = sum(w.numel() for w in words) * 4
    mask = -1 << bits
    ok = all(torch.equal(b, w & mask) for b, w in zip(backs, words))
    print(f"parts={nparts} {dt*1e3:.2f} ms  {tot/dt/1e9:.1f} GB/s ok={ok}")

for p in (1, 2, 3, 4):
    run(p)
