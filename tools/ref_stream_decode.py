"""Developer probe: speed of decoding REFERENCE-made containers (zlib streams, general inflater) on the GPU."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from datacompressionfloat_b200 import Codec
from oracle import oracle as O
from conftest import synth_words
codec = Codec.on_current_stream()
codec.set_profiling(True)
for kind, bits, nchunks in [("P", 0, 8), ("G", 8, 8)]:
    w = synth_words(kind, nchunks * 6291456 - 256)
    t0 = time.time(); cont = O.compress(w.view(np.uint8), bits); t_cpu = time.time() - t0
    t0 = time.time(); O.decompress(cont); t_cpu_d = time.time() - t0
    d = torch.from_numpy(cont).cuda()
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        back = codec.decompress(d)
        torch.cuda.synchronize(); dt = time.time() - t0
    ok = np.array_equal(back.cpu().numpy().view(np.uint8), O.erasebytes(w.view(np.uint8), bits))
    print(kind, bits, f"{w.size*4/2**20:.0f} MiB  gpu decode {dt*1e3:.1f} ms = {w.size*4/dt/1e9:.2f} GB/s  ok={ok}  stats={codec.stats()}  stage={ {k: round(v,1) for k,v in codec.stage_ms().items() if v>0.05} }  (cpu zlib: compress {t_cpu:.1f}s decompress {t_cpu_d:.1f}s)", flush=True)
