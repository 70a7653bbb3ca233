"""Developer probe: speed of decoding REFERENCE-made containers (zlib streams) on the GPU.
usage: python tools/ref_stream_decode.py [nchunks]   (chunks of 6291456 words; the CPU oracle makes the container)"""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from datacompressionfloat_b200 import Codec
from oracle import oracle as O
from conftest import synth_words
nchunks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
codec = Codec.on_current_stream()
codec.set_profiling(True)
for kind, bits in [("P", 0), ("G", 8), ("S", 12)]:
    w1 = synth_words(kind, 8 * 6291456 - 256)
    t0 = time.time(); c1 = O.compress(w1.view(np.uint8), bits); t_cpu = time.time() - t0
    t0 = time.time(); O.decompress(c1); t_cpu_d = time.time() - t0
    # tile the 8-chunk container: records are position independent
    reps = max(1, nchunks // 8)
    hdr = c1[:17].copy()
    hdr[:8] = np.frombuffer(np.uint64(w1.size * 4 * reps).tobytes(), np.uint8)
    cont = np.concatenate([hdr] + [c1[17:]] * reps)
    d = torch.from_numpy(cont).cuda()
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        back = codec.decompress(d)
        torch.cuda.synchronize(); dt = time.time() - t0
    gold = torch.from_numpy(O.erasebytes(w1.view(np.uint8), bits).view(np.int32)).cuda()
    ok = all(bool(torch.equal(back[r * w1.size:(r + 1) * w1.size].view(torch.int32), gold)) for r in range(reps))
    n = w1.size * reps
    print(kind, bits, f"{n*4/2**20:.0f} MiB  gpu decode {dt*1e3:.1f} ms = {n*4/dt/1e9:.2f} GB/s  ok={ok}  stats={codec.stats()}  stage={ {k: round(v,1) for k,v in codec.stage_ms().items() if v>0.05} }  (cpu zlib 1 thread per 192 MiB: compress {t_cpu:.1f}s decompress {t_cpu_d:.1f}s)", flush=True)
