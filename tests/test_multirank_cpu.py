"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: chunk-range sharding, the one exchange
of the path (all_gather of segment sizes -> exclusive scan -> container offsets) and assembly.
The per-rank compression is done by the oracle here (no GPU); the same logic drives bench.py with NCCL."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r})
sys.path.insert(0, {root!r} + "/tests")
from conftest import synth_words
from datacompressionfloat_b200 import chunk_range, segment_offsets, file_header
from oracle import oracle as O

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["MASTER_PORT"], rank=rank, world_size=world)
chk, bits = 4096, 7
w = synth_words("P", 11 * chk + 123)                      # every rank builds the same volume
nchunks = -(-w.size // chk)
lo, hi = chunk_range(nchunks, rank, world)
part = w[lo * chk: hi * chk]
# per-rank segment = chunk records only; rank 0 owns the header-exempt chunk (workers.c:90-94)
if rank == 0:
    seg = O.compress(part.view(np.uint8), bits, chk=chk)[17:]
else:
    # chunks that do not start the file are masked from word 0: prepend a dummy exempt chunk and drop its record
    pad = np.zeros(chk, np.uint32)
    c = O.compress(np.concatenate([pad, part]).view(np.uint8), bits, chk=chk)
    _, _, streams = O.parse_container(c)
    seg = c[streams[4]["offset"] - 16:]
sizes = torch.zeros(world, dtype=torch.int64)
dist.all_gather_into_tensor(sizes, torch.tensor([seg.size], dtype=torch.int64))   # the only exchange of the path
offs, total = segment_offsets(sizes.tolist())
out = np.memmap({out!r}, dtype=np.uint8, mode="r+")
if rank == 0:
    out[:17] = file_header(w.size * 4, chk)
out[offs[rank]: offs[rank] + seg.size] = seg
out.flush()
dist.barrier()
if rank == 0:
    whole = O.compress(w.view(np.uint8), bits, chk=chk)
    got = np.array(out[:total])
    assert total == whole.size and np.array_equal(got, whole), (total, whole.size)
    assert np.array_equal(O.decompress(got), O.erasebytes(w.view(np.uint8), bits))
    print("OK")
dist.destroy_process_group()
'''


def test_two_rank_sharded_container_matches_single():
    import subprocess
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "container.bin")
        np.zeros(1 << 20, np.uint8).tofile(out)
        script = os.path.join(d, "worker.py")
        Path(script).write_text(WORKER.format(root=str(ROOT), out=out))
        port = str(29500 + os.getpid() % 2000)
        procs = []
        for r in range(2):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
            procs.append(subprocess.Popen([sys.executable, script], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=240) for p in procs]
        for p, (so, se) in zip(procs, outs):
            assert p.returncode == 0, se[-2000:]
        assert "OK" in outs[0][0]
