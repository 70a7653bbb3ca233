#!/usr/bin/env python
"""bench.py -- float32 compress + decompress GB/s of the hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU implementation on the host cores

One step = one pass of the hot path over this rank's shard: compress (mask, byte-plane split,
per-plane deflate, container layout) followed by decompress (container walk, inflate, merge), all
device resident.  The volume is one synthetic float32 MRC file (1024-byte header + data):

    weak scaling (default)   N x 4 GiB of data: the 1024^3 volume of BASELINE.json configs[2] at N = 1
                             (1,073,742,080 words: 170 chunks + one of 4,194,560 words), 32 GiB at N = 8
    --scaling strong         --total-gib G (default 32, configs[4]) of data whatever N is

and every rank owns the contiguous chunk range SURVEY.md 8e gives it; the only cross-rank exchange is
the all_gather of the per-rank segment sizes that lays out the container.
`value` = uncompressed bytes of all ranks / max-over-ranks time per step, GB = 1e9 bytes.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "float32_compress_plus_decompress_throughput"
UNIT = "GB/s"
CHUNK_WORDS = 6 * 1048576
SUB = 16384
HDR_WORDS = 256
SEEDS = {"G": 1234, "P": 4321, "S": 7}
MATRIX = [("G", 0), ("G", 8), ("G", 16), ("P", 0), ("S", 12)]
REF_DECODE = [("G", 8), ("P", 0), ("S", 12)]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gib", type=float, default=4.0, help="GiB of float32 data per GPU (weak scaling; default 4 = 1024^3)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-gib", type=float, default=32.0, help="GiB of data of the whole volume with --scaling strong")
    ap.add_argument("--kind", default="G", choices=["G", "P", "S"], help="synthetic distribution (SURVEY 8d)")
    ap.add_argument("--bits", type=int, default=8, help="low mantissa bits erased")
    ap.add_argument("--batch-chunks", type=int, default=0, help="chunks per kernel batch (0 = library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-matrix", action="store_true", help="skip the G0/G8/G16/P0/S12 matrix and the reference-container decode")
    ap.add_argument("--matrix", action="store_true", help="run the matrix at N > 1 too")
    ap.add_argument("--cpu-chunks-per-file", type=int, default=6)
    return ap.parse_args()


# ----------------------------------------------------------------------------- the volume and its shards
def volume_words(a, world: int) -> int:
    gib = a.total_gib if a.scaling == "strong" else a.gib * world
    return int(gib * (1 << 30)) // 4 + HDR_WORDS


def shard_words(a, rank: int, world: int):
    """Contiguous chunk range of `rank` as a word range [lo, hi) of the volume."""
    W = volume_words(a, world)
    nchunks = -(-W // CHUNK_WORDS)
    per = -(-nchunks // world)
    lo = min(nchunks, rank * per) * CHUNK_WORDS
    hi = min(W, min(nchunks, (rank + 1) * per) * CHUNK_WORDS)
    return lo, max(lo, hi)


def gen_words(kind: str, nwords: int, seed_offset: int, device):
    """float32 data words on the device (torch generators; same distributions as synth.py)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEEDS[kind] + seed_offset)
    if kind == "G":
        d = torch.randn(nwords, generator=g, device=device, dtype=torch.float32)
    elif kind == "P":
        d = torch.poisson(torch.full((nwords,), 2.0, device=device, dtype=torch.float32), generator=g)
    else:
        x = torch.linspace(0, 4000 * np.pi, nwords, device=device, dtype=torch.float32)
        d = torch.sin(x) * torch.cos(x / 7) + 0.25 * torch.randn(nwords, generator=g, device=device, dtype=torch.float32)
    return d.view(torch.int32)


def workload_config(a, world, nwords_rank0):
    W = volume_words(a, world)
    full, rest = divmod(W, CHUNK_WORDS)
    vol = (f"{W:,} words = 1024-byte MRC header + {(W - HDR_WORDS) * 4 / 2**30:g} GiB float32 "
           f"({full} chunks of {CHUNK_WORDS:,} words + one of {rest:,})")
    return {"workload": f"mrc_full-style round trip of one synthetic MRC volume: {vol}; "
                        f"{'1024^3 (BASELINE configs[2])' if W == 1073742080 else ('32 GiB (BASELINE configs[4])' if W == 8589934848 else 'custom size')}, "
                        f"sharded by contiguous chunk range over {world} GPU(s)",
            "distribution": {"G": "normal(0,1)", "P": "poisson(2)", "S": "smooth+noise"}[a.kind], "mask_bits": a.bits,
            "chunk_words": CHUNK_WORDS, "volume_words": int(W), "words_per_gpu": int(nwords_rank0),
            "scaling_mode": a.scaling, "zlib_equivalent": "level 6, Z_RLE, raw deflate",
            "l2_policy": "inputs (>= 1 GiB per launch) exceed the 126 MB L2; no flush needed", "parallelism": f"chunk-range x{world}"}


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM clock and throttle reasons sampled every few ms by a thread (NVML) while the GPU is under load:
    from the first warm-up step to the end of the timed region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.timed = None   # (t0, t1) of the timed region, perf_counter
        self.err = None

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if visible:
                try:
                    idx = int(visible.split(",")[self.index])
                except (ValueError, IndexError):
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag:
                mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), mhz, rs))
                time.sleep(0.004)
        except Exception as ex:  # pragma: no cover
            self.err = repr(ex)

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"no samples ({self.err})"]}
        sm = [m for _, m, _ in self.samples]
        bits = 0
        for _, _, r in self.samples:
            bits |= r
        in_timed = [m for t, m, _ in self.samples if self.timed and self.timed[0] <= t <= self.timed[1]]
        return {"sm_mhz": float(np.median(in_timed if in_timed else sm)), "sm_max_mhz": float(getattr(self, "max_mhz", max(sm))),
                "samples": len(sm), "samples_in_timed_region": len(in_timed), "window": "first warm-up step .. end of the timed region",
                "reasons": sorted(v for k, v in self.REASONS.items() if bits & k)}


# ----------------------------------------------------------------------------- CPU baseline: the reference on the host cores
def host_sample_fn(kind: str):
    """Sample generator for the CPU legs (numpy / torch-CPU, no GPU needed)."""
    import torch

    def fn(nwords: int, i: int) -> np.ndarray:
        g = torch.Generator()
        g.manual_seed(SEEDS[kind] + 1000 + i)
        if kind == "G":
            d = torch.randn(nwords, generator=g, dtype=torch.float32)
        elif kind == "P":
            d = torch.poisson(torch.full((nwords,), 2.0), generator=g)
        else:
            x = torch.linspace(0, 4000 * np.pi, nwords)
            d = torch.sin(x) * torch.cos(x / 7) + 0.25 * torch.randn(nwords, generator=g)
        w = d.numpy().view(np.uint32).copy()
        w[:HDR_WORDS] = 0
        return w
    return fn


class RefFiles:
    """`cores` sample files of `chunks_per_file` chunks in /dev/shm, zipped / unzipped by the UNMODIFIED reference
    (oracle/_ref/mrc_tarx_c -n cores: one file per pthread worker, reference src/main/mrc_tarx.c:134-176)."""

    def __init__(self, kind, bits, chunks_per_file, cores=None):
        from oracle import oracle as O
        self.O = O
        self.kind, self.bits = kind, bits
        self.cores = cores or (os.cpu_count() or 1)
        self.words_per_file = chunks_per_file * CHUNK_WORDS
        self.chunks_per_file = chunks_per_file
        self.tmp = Path(tempfile.mkdtemp(prefix="mrcz_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
        self.src_dir, self.zip_dir, self.out_dir = self.tmp / "src", self.tmp / "zip", self.tmp / "out"
        for d in (self.src_dir, self.zip_dir, self.out_dir):
            d.mkdir()
        fn = host_sample_fn(kind)
        self.names = []
        for i in range(self.cores):
            p = self.src_dir / f"v{i:03d}.mrc"
            fn(self.words_per_file, i).tofile(p)
            self.names.append(p)
        self.total = self.cores * self.words_per_file * 4
        (self.tmp / "zip.txt").write_text("".join(f"{p}\n" for p in self.names))
        (self.tmp / "unzip.txt").write_text("".join(f"{self.zip_dir / (p.name + '.zip')}\n" for p in self.names))
        self.exe = str(O.REF_DIR / "mrc_tarx_c")

    def zip(self):
        t0 = time.perf_counter()
        subprocess.run([self.exe, "-i", str(self.tmp / "zip.txt"), "-t", "zip", "-o", str(self.zip_dir), "-b", str(self.bits),
                        "-n", str(self.cores), "-d", "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return time.perf_counter() - t0

    def unzip(self):
        t0 = time.perf_counter()
        subprocess.run([self.exe, "-i", str(self.tmp / "unzip.txt"), "-t", "unzip", "-o", str(self.out_dir),
                        "-n", str(self.cores), "-d", "1"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return time.perf_counter() - t0

    def zsize(self):
        return sum((self.zip_dir / (p.name + ".zip")).stat().st_size - 17 for p in self.names)

    def container(self):
        """One container out of the files' chunk records (chunks are independent): what a reference user's archive of
        cores x chunks_per_file chunks looks like to the decoder.  -> (uint8 array, words array of the originals)."""
        recs = [np.fromfile(self.zip_dir / (p.name + ".zip"), dtype=np.uint8)[17:] for p in self.names]
        nwords = self.cores * self.words_per_file
        hdr = np.zeros(17, dtype=np.uint8)
        hdr[:8] = np.frombuffer(np.uint64(nwords * 4).tobytes(), dtype=np.uint8)
        hdr[8:12] = np.frombuffer(np.uint32(CHUNK_WORDS).tobytes(), dtype=np.uint8)
        words = np.concatenate([np.fromfile(p, dtype=np.uint32) for p in self.names])
        return np.concatenate([hdr] + recs), words

    def close(self):
        shutil.rmtree(self.tmp, ignore_errors=True)


def cpu_reference_run(kind: str, bits: int, chunks_per_file: int, steps: int, warmup: int, keep=None):
    """Times the reference's own multithreaded path.  Falls back to the single-threaded oracle port when oracle/_ref
    is not there.  Returns dict(value GB/s, cores, kind, sample, ...).  keep: a list that receives the RefFiles
    (zipped) instead of deleting them."""
    from oracle import oracle as O
    if O.have_ref():
        rf = RefFiles(kind, bits, chunks_per_file)
        try:
            for _ in range(warmup):
                rf.zip(); rf.unzip()
            tz, tu = [], []
            for _ in range(steps):
                tz.append(rf.zip()); tu.append(rf.unzip())
            tzm, tum = float(np.mean(tz)), float(np.mean(tu))
            total, cores = rf.total, rf.cores
            return dict(value=total / (tzm + tum) / 1e9, unit=UNIT, cores=cores, kind="reference",
                        sample=f"{cores} files x {chunks_per_file} chunks ({total / 2**30:.2f} GiB {kind} b={bits}) in /dev/shm, "
                               f"mrc_tarx_c -n {cores} zip (-d 0) + unzip (-d 1), reference built -O2 with its own zlib 1.2.8",
                        compress_GBs=total / tzm / 1e9, decompress_GBs=total / tum / 1e9, ratio=rf.zsize() / total,
                        ms_per_step=(tzm + tum) * 1e3, bytes=total)
        finally:
            if keep is not None:
                keep.append(rf)
            else:
                rf.close()
    # port: the oracle restatement, one thread
    w = host_sample_fn(kind)(min(chunks_per_file, 3) * CHUNK_WORDS, 0)
    total = w.size * 4
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        c = O.compress(w.view(np.uint8), bits)
        O.decompress(c)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts))
    return dict(value=total / t / 1e9, unit=UNIT, cores=1, kind="port",
                sample=f"{total / 2**20:.0f} MiB {kind} b={bits}, oracle C port (system zlib), 1 thread",
                ratio=(c.size - 17) / total, ms_per_step=t * 1e3, bytes=total)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    r = cpu_reference_run(a.kind, a.bits, a.cpu_chunks_per_file, a.steps, a.warmup)
    lo, hi = shard_words(a, 0, max(world, a.gpus))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(a, max(world, a.gpus), hi - lo),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "compress_GBs": r.get("compress_GBs"), "decompress_GBs": r.get("decompress_GBs"), "ratio": r.get("ratio"),
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def bind_near_gpu(local_rank):
    """Multi-rank runs: keep this process (and the pinned host buffers it is about to allocate) on the NUMA node
    of its GPU, so that eight ranks do not pull 100+ GB per step across the socket interconnect.  When sysfs does not
    know the GPU's node (-1 on virtualised hosts) the ranks are spread evenly over the nodes there are."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        how = "sysfs"
        nodes = sorted(int(d.name[4:]) for d in Path("/sys/devices/system/node").glob("node[0-9]*"))
        if node < 0:
            if len(nodes) < 2:
                return {"numa_node": None, "nodes": len(nodes)}
            world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
            node = nodes[(local_rank * len(nodes)) // max(world, 1)]
            how = "spread (sysfs says -1)"
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus), "how": how, "nodes": len(nodes)}
    except Exception as e:  # no sysfs / unknown properties: run unbound
        return {"numa_node": None, "note": type(e).__name__}


# ----------------------------------------------------------------------------- the B200 arm
def touched_bytes(nb, out_bytes, cs, bits):
    """Bytes each stage has to read + write on THIS input (derived from the call's counters), next to the algorithmic
    bytes of SURVEY 8d: planes the mask erased, RAW streams and stored sub-blocks are not entropy coded."""
    streams = max(cs["streams"], 1)
    total_sub = int(np.ceil(nb / 4 / SUB)) * 4 if nb else 0
    zero_sub = cs.get("zero_subblocks", 0)
    stored_sub = cs["stored_subblocks"]
    coded_sub = max(total_sub - zero_sub - stored_sub, 0)
    raw_bytes = cs["raw_streams"] * (nb / streams)
    known_zero = bits >= 8
    coded_out = max(out_bytes - 16 * cs["chunks"] - stored_sub * (SUB + 15), 0)
    payload_scanned = max(out_bytes - raw_bytes, 0)
    stored_nonraw = max(stored_sub * SUB - raw_bytes, 0)
    return {
        "split": nb + nb * (4 - min(bits // 8, 4)) / 4,   # byte planes the mask erases are not written (nor read again)
        "encode": stored_sub * 2048 + coded_sub * (2048 + SUB / 4 + SUB) + zero_sub * (0 if known_zero else 2048 + SUB) + coded_out,
        "gather": 2 * out_bytes,
        "markers": payload_scanned,
        "inflate_fast": payload_scanned + coded_sub * SUB + stored_nonraw,
        "merge": nb + (total_sub - zero_sub) * SUB,
    }


def run_b200(a):
    import torch
    import torch.distributed as dist
    from datacompressionfloat_b200 import Codec, lib as mzlib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world != 1:
        print(f"warning: WORLD_SIZE {world} != --gpus {a.gpus}", file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device: there is no CPU fallback")
    affinity = bind_near_gpu(local) if world > 1 else {"numa_node": None}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's version / debug lines must not land on stdout next to the JSON line (NCCL_DEBUG=VERSION printf's there)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # this rank's contiguous chunk range of the volume; rank 0's starts with the 1024-byte MRC header (exempt from masking)
    lo, hi = shard_words(a, rank, world)
    nwords = hi - lo
    exempt = HDR_WORDS if rank == 0 else 0
    words = torch.empty(max(nwords, 4), dtype=torch.int32, device=dev)[:nwords]

    def load(kind):
        words[exempt:] = gen_words(kind, nwords - exempt, rank, dev)
        if rank == 0:
            words[:HDR_WORDS] = 0
            words[0:3] = 1024
            words[3] = 2

    load(a.kind)
    codec = Codec.on_current_stream(batch_chunks=a.batch_chunks or None)
    codec.set_profiling(True)
    cont_buf = torch.empty(Codec.compress_bound(nwords), dtype=torch.uint8, device=dev)
    out_words = torch.empty(max(nwords, 4), dtype=torch.int32, device=dev)
    sizes_all = torch.zeros(world, dtype=torch.int64, device=dev)
    my_size = torch.zeros(1, dtype=torch.int64, device=dev)
    state = {}

    def step(bits, timed=None):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed is not None else None
        if e: e[0].record()
        seg = codec.compress(words, bits, exempt_words=exempt, write_file_header=False, out=cont_buf)
        cs, cms = codec.stats(), codec.stage_ms()
        if world > 1:  # the one exchange of the path: segment sizes -> container offsets (exclusive scan)
            my_size[0] = seg.numel()
            dist.all_gather_into_tensor(sizes_all, my_size)
        if e: e[1].record()
        back = codec.decompress(seg, has_file_header=False, nwords=nwords, out=out_words)
        ds, dms = codec.stats(), codec.stage_ms()
        if e:
            e[2].record()
            timed.append((e, cms, dms))
        state.update(seg=seg, back=back, cs=cs, ds=ds)

    def check(bits):
        mask = (-1 << bits) if bits < 32 else 0
        ref = words.clone()
        ref[exempt:] &= mask
        ok = bool(torch.equal(ref, state["back"]))
        del ref
        return ok

    def measure(bits, warmup, steps):
        """W untimed + K timed steps on the data now in `words`; max over ranks; bit-exactness checked first."""
        for _ in range(warmup):
            step(bits)
        if not check(bits):
            raise SystemExit("round trip is not bit-exact: refusing to report a number")
        timed = []
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        t0.record()
        for _ in range(steps):
            step(bits, timed)
        t1.record()
        barrier()
        wall1 = time.perf_counter()
        elapsed_ms = t0.elapsed_time(t1)
        comp_ms = float(np.mean([e[0].elapsed_time(e[1]) for e, _, _ in timed]))
        decomp_ms = float(np.mean([e[1].elapsed_time(e[2]) for e, _, _ in timed]))
        stage_c = {k: float(np.mean([c[k] for _, c, _ in timed])) for k in timed[0][1]}
        stage_d = {k: float(np.mean([d[k] for _, _, d in timed])) for k in timed[0][2]}
        agg = torch.tensor([elapsed_ms, comp_ms, decomp_ms], dtype=torch.float64, device=dev)
        tot = torch.tensor([float(nwords * 4), float(state["seg"].numel())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(agg, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        elapsed_ms, comp_ms, decomp_ms = [float(x) for x in agg.tolist()]
        total_bytes, total_comp = [float(x) for x in tot.tolist()]
        ms_per_step = elapsed_ms / steps
        stages = {k: stage_c.get(k, 0.0) + stage_d.get(k, 0.0) for k in set(stage_c) | set(stage_d)}
        return dict(ms_per_step=ms_per_step, value=total_bytes / (ms_per_step * 1e-3) / 1e9, comp_ms=comp_ms, decomp_ms=decomp_ms,
                    total_bytes=total_bytes, total_comp=total_comp, stages=stages, wall=(wall0, wall1),
                    cs=dict(state["cs"]), ds=dict(state["ds"]), seg_bytes=int(state["seg"].numel()))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # samples clocks / throttle reasons from the warm-up through the timed region
    head = measure(a.bits, a.warmup, a.steps)
    sampler.timed = head["wall"]
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step, value = head["ms_per_step"], head["value"]
    total_bytes, total_comp = head["total_bytes"], head["total_comp"]
    stages = head["stages"]

    # ---- end to end through the C ABI with HOST buffers (pinned), H2D and D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        mask = (-1 << a.bits) if a.bits < 32 else 0
        h_in = torch.empty(nwords, dtype=torch.int32).pin_memory()
        h_in.copy_(words)
        cap = Codec.compress_bound(nwords)
        h_cont = torch.empty(cap, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(nwords, dtype=torch.int32).pin_memory()
        h_ref = (words[exempt:exempt + 65536] & mask).cpu().numpy()
        torch.cuda.synchronize()

        def e2e_step():
            sz = codec.compress_host_ptr(h_in.data_ptr(), nwords, a.bits, h_cont.data_ptr(), cap, exempt_words=exempt,
                                         write_file_header=False)
            got = codec.decompress_host_ptr(h_cont.data_ptr(), sz, h_out.data_ptr(), nwords, has_file_header=False, nwords=nwords)
            return sz, got

        for _ in range(max(1, min(a.warmup, 2))):
            sz, got = e2e_step()
        barrier()
        w0 = time.perf_counter()
        k2 = max(1, min(a.steps, 3))
        for _ in range(k2):
            sz, got = e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - w0) / k2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_ok = bool(got == nwords and np.array_equal(h_out.numpy()[exempt:exempt + 65536], h_ref) and
                      np.array_equal(h_out.numpy()[-4096:], (words[-4096:] & mask).cpu().numpy()))
        e2e = {"value": total_bytes / float(dt) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(nwords * 4 + sz),
               "d2h_bytes_per_step": int(sz + nwords * 4), "steps": k2, "ok": e2e_ok,
               "api": "mzb_compress_host + mzb_decompress_host on pinned host buffers"}

        # the ceiling of that leg: the same bytes over the same pinned buffers with no kernel in between, all ranks at
        # once -- compress moves nb up while sz comes down, decompress sz up while nb comes down
        s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        d_a, d_b = words.view(torch.uint8), cont_buf
        nbytes = nwords * 4

        def copy_step():
            with torch.cuda.stream(s_up):
                d_a.copy_(h_in.view(torch.uint8), non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_cont[:sz].copy_(d_b[:sz], non_blocking=True)
            s_up.synchronize(); s_dn.synchronize()
            with torch.cuda.stream(s_up):
                d_b[:sz].copy_(h_cont[:sz], non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_out.view(torch.uint8).copy_(out_words.view(torch.uint8)[:nbytes], non_blocking=True)
            s_up.synchronize(); s_dn.synchronize()

        copy_step()
        barrier()
        w0 = time.perf_counter()
        for _ in range(k2):
            copy_step()
        barrier()
        dtc = torch.tensor([(time.perf_counter() - w0) / k2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
        e2e["ceiling_GBs"] = total_bytes / float(dtc) / 1e9
        e2e["frac_of_ceiling"] = e2e["value"] / e2e["ceiling_GBs"]
        e2e["ceiling_note"] = ("pinned H2D + D2H of the step's bytes on two streams, all ranks at once, no kernels: "
                               "what the host memory / PCIe path of this box gives the pipeline")
        load(a.kind)   # the probe overwrote the device copy of the volume
        del h_cont, h_out

        # ---- the same round trip at the drop-in boundary proper: zip_compress / zip_uncompress (reference adapt.h:30-31)
        # on files in /dev/shm, with the -d semantics of the reference arm (zip writes its output, unzip is -d 1)
        e2e_file = None
        if world == 1 and os.path.isdir("/dev/shm"):
            tmpd = Path(tempfile.mkdtemp(prefix="mrcz_e2e_", dir="/dev/shm"))
            try:
                L = mzlib.load()
                src, dst, back = tmpd / "v.mrc", tmpd / "v.mrc.zip", tmpd / "v.out"
                h_in.numpy().tofile(src)
                flag = ctypes.c_int.in_dll(L, "isTestThroughput")
                ctx = mzlib.CtxT()

                def file_step():
                    L.init_context(ctypes.byref(ctx))
                    flag.value = 0
                    t0 = time.perf_counter()
                    rc1 = L.zip_compress(ctypes.byref(ctx), str(src).encode(), str(dst).encode(), a.bits)
                    t1 = time.perf_counter()
                    flag.value = 1
                    rc2 = L.zip_uncompress(ctypes.byref(ctx), str(dst).encode(), str(back).encode())
                    t2 = time.perf_counter()
                    flag.value = 0
                    return rc1, rc2, t1 - t0, t2 - t1

                file_step()
                rs = [file_step() for _ in range(2)]
                tz = float(np.mean([r[2] for r in rs])); tu = float(np.mean([r[3] for r in rs]))
                # correctness of the files (one more unzip that does write)
                L.init_context(ctypes.byref(ctx))
                L.zip_uncompress(ctypes.byref(ctx), str(dst).encode(), str(back).encode())
                got = np.fromfile(back, dtype=np.int32)
                okf = bool(all(r[0] == 0 and r[1] == 0 for r in rs) and got.size == nwords and
                           np.array_equal(got[exempt:exempt + 65536], h_ref) and
                           np.array_equal(got[-4096:], (words[-4096:] & mask).cpu().numpy()))
                e2e_file = {"value": nwords * 4 / (tz + tu) / 1e9, "unit": UNIT, "zip_GBs": nwords * 4 / tz / 1e9,
                            "unzip_GBs": nwords * 4 / tu / 1e9, "ok": okf, "container_bytes": dst.stat().st_size,
                            "api": "zip_compress (-d 0) + zip_uncompress (-d 1) of libmrczip_b200.so on files in /dev/shm: "
                                   "the reference arm's flags, one file instead of one file per core"}
            except Exception as ex:
                e2e_file = {"value": None, "unit": UNIT, "error": repr(ex)}
            finally:
                shutil.rmtree(tmpd, ignore_errors=True)
        del h_in

    # ---- the five inputs SURVEY 8d names, same volume size, one warm step + 5 timed each
    matrix = None
    kept = []
    cpu = None
    if (world == 1 or a.matrix) and not a.no_matrix:
        matrix = {}
        for kind, bits in MATRIX:
            if kind == a.kind and bits == a.bits:
                m = head
            else:
                load(kind)
                m = measure(bits, 1, 5)
            matrix[f"{kind}{bits}"] = {"value": m["value"], "ms_per_step": m["ms_per_step"],
                                       "encode_ms": m["stages"].get("encode", 0.0), "inflate_ms": m["stages"].get("inflate_fast", 0.0),
                                       "compress_GBs": m["total_bytes"] / (m["comp_ms"] * 1e-3) / 1e9,
                                       "decompress_GBs": m["total_bytes"] / (m["decomp_ms"] * 1e-3) / 1e9,
                                       "ratio": m["total_comp"] / m["total_bytes"],
                                       "raw_streams": m["cs"]["raw_streams"], "stored_subblocks": m["cs"]["stored_subblocks"]}
        load(a.kind)

    if rank == 0 and not a.no_cpu_baseline and world == 1:
        try:
            r = cpu_reference_run(a.kind, a.bits, a.cpu_chunks_per_file, 1, 0, keep=kept)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu.update(compress_GBs=r.get("compress_GBs"), decompress_GBs=r.get("decompress_GBs"), ratio=r.get("ratio"))
        except Exception as ex:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    # ---- containers written by the reference itself (zlib streams, no sub-block framing): decode on the GPU
    if matrix is not None and world == 1:
        refdec = {}
        try:
            from oracle import oracle as O
            if O.have_ref():
                for kind, bits in REF_DECODE:
                    rf = next((k for k in kept if k.kind == kind and k.bits == bits), None)
                    own = rf is None
                    if own:
                        rf = RefFiles(kind, bits, a.cpu_chunks_per_file)
                        rf.zip()
                    try:
                        cont, orig = rf.container()
                        d_cont = torch.from_numpy(cont).to(dev)
                        nw = orig.size
                        d_out = out_words[:nw] if nw <= out_words.numel() else torch.empty(nw, dtype=torch.int32, device=dev)
                        for _ in range(2):
                            back = codec.decompress(d_cont, out=d_out)
                        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                        evs[0].record()
                        for _ in range(3):
                            back = codec.decompress(d_cont, out=d_out)
                        evs[1].record()
                        torch.cuda.synchronize()
                        ms = evs[0].elapsed_time(evs[1]) / 3
                        st = codec.stats()
                        exp = orig & np.uint32((0xFFFFFFFF << bits) & 0xFFFFFFFF if bits < 32 else 0)
                        wpf = rf.words_per_file
                        for i in range(rf.cores):   # every file kept its own 256 header words unmasked
                            exp[i * wpf:i * wpf + HDR_WORDS] = orig[i * wpf:i * wpf + HDR_WORDS]
                        okd = bool(np.array_equal(back.cpu().numpy().view(np.uint32), exp))
                        refdec[f"{kind}{bits}"] = {"GBs": nw * 4 / (ms * 1e-3) / 1e9, "ms": ms, "bytes": int(nw * 4), "chunks": int(nw // CHUNK_WORDS),
                                                   "bit_exact": okd, "blockpar_streams": st["blockpar_streams"],
                                                   "general_streams": st["general_streams"]}
                        del d_cont
                    finally:
                        if own:
                            rf.close()
        except Exception as ex:
            refdec["error"] = repr(ex)
        matrix["ref_container_decode_GBs"] = refdec
    for k in kept:
        k.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (stage), live CUDA-event durations from the timed region
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ratio = head["seg_bytes"] / (nwords * 4.0)
    nb = nwords * 4.0
    alg = {  # algorithmic bytes per launch group, r = compressed / original
        "split": 2 * nb, "merge": 2 * nb, "encode": nb + nb * ratio, "gather": 2 * nb * ratio,
        "inflate_fast": nb * ratio + nb, "rawcopy": 2 * nb * ratio, "markers": nb * ratio,
    }
    touched = touched_bytes(nb, head["seg_bytes"], head["cs"], a.bits)
    kernels = []
    for k, ms in sorted(stages.items(), key=lambda kv: -kv[1]):
        if ms <= 0 or k not in alg:
            continue
        ach = alg[k] / (ms * 1e-3) / 1e9
        row = {"kernel": k, "ms_per_step": ms, "achieved": ach, "frac": ach / peak, "algorithmic_bytes": alg[k]}
        if k in touched:
            row["touched_bytes"] = touched[k]
            row["frac_touched"] = touched[k] / (ms * 1e-3) / 1e9 / peak
        kernels.append(row)
    dom = kernels[0] if kernels else None
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and dom:
        try:
            tj = json.loads(tf.read_text())
            traffic = tj.get(dom["kernel"])
            traffic_src = tj.get("_source")
        except Exception:
            traffic = None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac"], "frac_touched": dom.get("frac_touched"), "touched_bytes": dom.get("touched_bytes"),
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "note": "achieved = algorithmic bytes of the stage per step (SURVEY 8d) / its CUDA-event time (all launches of the "
                            "stage in a step); frac_touched = bytes the stage must read + write on THIS input (planes the mask erased, "
                            "RAW streams and stored sub-blocks are not coded) / time / peak"}

    # whole-pipeline view: bytes a perfectly fused implementation would have to move (SURVEY 8d lower bound,
    # 4 + 4r per word each way) over the measured step time, as a fraction of the measured HBM peak
    fused_bytes = 2 * (nb + nb * ratio) * world
    pipeline = {"fused_lower_bound_bytes": fused_bytes, "achieved": fused_bytes / (ms_per_step * 1e-3) / 1e9 / world,
                "unit": "GB/s per GPU", "frac_of_hbm_peak": fused_bytes / (ms_per_step * 1e-3) / 1e9 / world / peak}
    launches = (head["cs"]["kernel_launches"] + head["ds"]["kernel_launches"]) * a.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(a, world, nwords),
        "compress_GBs": total_bytes / (head["comp_ms"] * 1e-3) / 1e9, "decompress_GBs": total_bytes / (head["decomp_ms"] * 1e-3) / 1e9,
        "ratio": total_comp / total_bytes, "ratio_definition": "compressed/original (reference zip.c:434), chunk records only",
        "bit_exact_roundtrip": True, "roofline": roofline, "roofline_kernels": kernels, "pipeline_roofline": pipeline,
        "cpu_baseline": cpu, "e2e": e2e, "e2e_file": (e2e_file if not a.no_e2e else None), "matrix": matrix,
        "gpu_launches": int(launches), "clocks": clocks,
        "decode_stats": {k: head["ds"][k] for k in ("general_streams", "fast_failed")},
        "host_affinity": affinity,
        "encode_stats": {k: head["cs"][k] for k in ("raw_streams", "stored_subblocks", "zero_subblocks", "streams") if k in head["cs"]},
        "stage_ms": {k: round(v, 4) for k, v in stages.items() if v > 0},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(line: dict) -> None:
    """The one JSON line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    a = parse_args()
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner, the reference
    # binaries' tables) is sent to stderr instead
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
