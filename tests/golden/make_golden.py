"""Generate the golden fixtures from the REFERENCE ITSELF (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run here (the container with /root/reference); the fixtures travel to the GPU box.

For every case:  input file, the reference's container (mrc_tar_c -t zip, its own zlib 1.2.8),
the reference's erasebytes output, and the four byte planes from the reference's
split_float_to_byte_stream (via oracle/ref_harness.c).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import synth_words  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = Path(__file__).resolve().parent

CASES = [
    # name, kind, data words, extra tail bytes, bits
    ("G_b0", "G", 9000, 0, 0),
    ("G_b8_ragged", "G", 7001, 3, 8),
    ("P_b4", "P", 12000, 0, 4),
    ("S_b12", "S", 8192, 1, 12),
    ("Z_b31", "Z", 5000, 0, 31),
    ("R_b32", "R", 3000, 2, 32),
    ("tiny_header_only", "G", 0, 0, 5),       # exactly the 1024-byte MRC header
    ("sub_header", "R", 0, 0, 9),             # cut below to 100 words: shorter than the exempt header
]


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    manifest = {"reference_zlib": O.ref_zlib_version(), "cases": []}
    for name, kind, n, tail, bits in CASES:
        w = synth_words(kind, n, seed=len(name))
        if name == "sub_header":
            w = w[:100]
        raw = np.concatenate([w.view(np.uint8), np.arange(tail, dtype=np.uint8)])
        files = {"input": f"{name}.in", "ref_container": f"{name}.zip", "ref_erasebytes": f"{name}.erase",
                 "ref_planes": [f"{name}.p{j}" for j in range(4)]}
        raw.tofile(HERE / files["input"])
        O.ref_compress(raw, bits).tofile(HERE / files["ref_container"])
        O.ref_erasebytes(raw, bits).tofile(HERE / files["ref_erasebytes"])
        _, planes = O.ref_split(raw, bits)
        for j in range(4):
            planes[j].tofile(HERE / files["ref_planes"][j])
        manifest["cases"].append(dict(name=name, bits=bits, **files))
        print(name, raw.size, "->", (HERE / files["ref_container"]).stat().st_size)
    (HERE / "manifest.json").write_text(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
