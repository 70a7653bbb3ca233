"""CPU study for DESIGN 9 lead #1: what would it cost in size to cut the lane pieces of a sub-block from 512 bytes
(today: lane l owns one contiguous piece, so the emit kernel needs a counting pass for the per-lane bit offsets) to
64 bytes (window-interleaved pieces: a warp scan inside the window would replace the counting pass)?

The tokeniser restarts at every piece start (no match reaches back over it, the first 6 repeats of a run are literals),
so shorter pieces mean shorter runs.  Cost model: tokens of a group of 32 sub-blocks (512 KiB) coded with the group's
own static entropy (what the group Huffman code approaches) + extra bits + 1 distance bit per match.  Pure numpy,
no GPU, no library calls."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datacompressionfloat_b200 import synth

SUB, GROUP, HOLD = 16384, 32, 6
LEN_BASE = [3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258]
LEN_EXTRA = [0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0]

def len_code(l):
    for c in range(28, -1, -1):
        if l >= LEN_BASE[c]:
            return c
    raise ValueError(l)

def tokens_of_piece(b, prev):
    """(literal byte histogram[256], list of match lengths) of one piece under the hold-after-6 rule"""
    n = b.size
    eq = np.empty(n, bool)
    eq[0] = prev is not None and b[0] == prev
    eq[1:] = b[1:] == b[:-1]
    # run id: starts where eq is False
    starts = np.flatnonzero(~eq)
    if starts.size == 0 or starts[0] != 0:
        starts = np.concatenate([[0], starts])
    ends = np.concatenate([starts[1:], [n]])
    rl = ends - starts                         # run lengths (first byte of a run is a literal: not a repeat)
    first_is_repeat = np.zeros(rl.size, bool)
    first_is_repeat[0] = eq[0]
    repeats = rl - 1 + first_is_repeat         # repeats of the run's byte inside this piece
    held = np.maximum(repeats - HOLD, 0)       # withheld bytes
    lit_counts = rl - held
    hist = np.zeros(256, np.int64)
    np.add.at(hist, b[starts], lit_counts)
    matches = []
    for h, v in zip(held[held > 0], b[starts][held > 0]):
        h = int(h)
        while h >= 258:
            matches.append(258); h -= 258
        if h >= 3: matches.append(h)
        elif h: hist[v] += h
    return hist, matches

def group_cost_bits(plane_bytes, piece):
    total = 0.0
    for g0 in range(0, plane_bytes.size, SUB * GROUP):
        grp = plane_bytes[g0:g0 + SUB * GROUP]
        hist = np.zeros(286, np.int64)
        extra = 0
        nsub = (grp.size + SUB - 1) // SUB
        for s0 in range(0, grp.size, SUB):
            sub = grp[s0:s0 + SUB]
            for p0 in range(0, sub.size, piece):
                h, ms = tokens_of_piece(sub[p0:p0 + piece], sub[p0 - 1] if p0 else None)
                hist[:256] += h
                for m in ms:
                    c = len_code(m); hist[257 + c] += 1; extra += LEN_EXTRA[c] + 1
        hist[256] = nsub
        f = hist[hist > 0].astype(np.float64)
        total += float(-(f * np.log2(f / f.sum())).sum()) + extra + nsub * (80 * 8 + 40)   # + header & marker per sub-block
    return total

def planes_of(kind, bits, nwords):
    w = synth.volume_data(kind, (1, 1, nwords)).view(np.uint32) & np.uint32((0xFFFFFFFF << bits) & 0xFFFFFFFF)
    return [((w >> (8 * j)) & 0xFF).astype(np.uint8) for j in range(4)]

if __name__ == "__main__":
    n = 1 << 20   # 1 Mi words = 2 groups per plane
    print("| input | plane | bytes/byte today (512 B pieces) | with 64 B pieces | change |\n|---|---|---|---|---|")
    for kind, bits in [("G", 8), ("P", 0), ("S", 12)]:
        for j, pl in enumerate(planes_of(kind, bits, n)):
            a = group_cost_bits(pl, 512) / 8 / pl.size
            if a > 0.97:
                continue   # stored / RAW anyway
            b = group_cost_bits(pl, 64) / 8 / pl.size
            print(f"| {kind} b={bits} | {j} | {a:.4f} | {b:.4f} | {100 * (b - a) / a:+.1f} % ({100 * (b - a):+.2f} % of the plane) |")
