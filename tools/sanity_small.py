"""Small, fast exercise of every kernel path (used under compute-sanitizer on the GPU box)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from conftest import synth_words  # noqa: E402
from datacompressionfloat_b200 import Codec  # noqa: E402
from oracle import oracle as O  # noqa: E402

codec = Codec(0)
for kind, bits, n, chk in [("G", 8, 70001, 16384), ("P", 0, 50000, 65536), ("R", 0, 40003, 16384), ("Z", 3, 33000, 1000),
                           ("S", 12, 20000, 4096), ("G", 16, 3, 16384), ("S", 12, 2300000, 1 << 21), ("P", 0, 2200000, 1 << 21)]:
    w = synth_words(kind, n)[: max(n, 1)]
    d = torch.from_numpy(w.view(np.int32)).cuda()
    cont = codec.compress(d, bits, chk=chk)
    back = codec.decompress(cont)
    gold = O.erasebytes(w.view(np.uint8), bits)
    assert np.array_equal(back.cpu().numpy().view(np.uint8), gold), (kind, bits)
    codec.set_inflate_variant(1)          # the full group inflater alone (the lean kernel's fallback)
    assert np.array_equal(codec.decompress(cont).cpu().numpy().view(np.uint8), gold), (kind, bits, "full")
    codec.set_inflate_variant(0)
    assert np.array_equal(O.decompress(cont.cpu().numpy()), gold)
    ref = O.compress(w.view(np.uint8), bits, chk=chk)
    back2 = codec.decompress(torch.from_numpy(ref).cuda())
    assert np.array_equal(back2.cpu().numpy().view(np.uint8), gold)
    h = codec.compress_host(w, bits, chk=chk)
    assert np.array_equal(codec.decompress_host(h).view(np.uint8), gold)
    print("ok", kind, bits, n, chk, cont.numel(), flush=True)
codec.close()
print("sanity_small: all ok")
