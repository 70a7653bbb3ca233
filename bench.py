#!/usr/bin/env python
"""bench.py -- float32 compress + decompress GB/s of the hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU implementation on the host cores

One step = one pass of the hot path over this rank's shard: compress (mask, byte-plane split,
per-plane deflate, container layout) followed by decompress (container walk, inflate, merge), all
device resident.  Weak scaling: every rank owns a contiguous chunk range of 4 GiB of one (N x 4 GiB)
synthetic float32 volume (1024^3 per GPU; 32 GiB at N = 8, BASELINE.json configs[2] and [4]); the
only cross-rank exchange is the all_gather of the per-rank segment sizes that lays out the container.
`value` = uncompressed bytes of all ranks / max-over-ranks time per step, GB = 1e9 bytes.
"""
from __future__ import annotations

import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gib", type=float, default=4.0, help="GiB of float32 per GPU (default 4 = 1024^3)")
    ap.add_argument("--kind", default="G", choices=["G", "P", "S"], help="synthetic distribution (SURVEY 8d)")
    ap.add_argument("--bits", type=int, default=8, help="low mantissa bits erased")
    ap.add_argument("--batch-chunks", type=int, default=0, help="chunks per kernel batch (0 = library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-chunks-per-file", type=int, default=6)
    return ap.parse_args()


# ----------------------------------------------------------------------------- synthetic data
def gen_words(kind: str, nwords: int, seed_offset: int, device):
    """float32 volume shard on the device (torch generators; same distributions as synth.py)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed({"G": 1234, "P": 4321, "S": 7}[kind] + seed_offset)
    if kind == "G":
        d = torch.randn(nwords, generator=g, device=device, dtype=torch.float32)
    elif kind == "P":
        d = torch.poisson(torch.full((nwords,), 2.0, device=device, dtype=torch.float32), generator=g)
    else:
        x = torch.linspace(0, 4000 * np.pi, nwords, device=device, dtype=torch.float32)
        d = torch.sin(x) * torch.cos(x / 7) + 0.25 * torch.randn(nwords, generator=g, device=device, dtype=torch.float32)
    return d.view(torch.int32)


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
def cpu_reference_run(kind: str, bits: int, chunks_per_file: int, steps: int, warmup: int, sample_words_fn):
    """Times the reference's own multithreaded path: mrc_tarx_c -n T (one file per pthread worker,
    reference src/main/mrc_tarx.c:134-176) on T files in /dev/shm.  Falls back to the single-threaded
    oracle port when oracle/_ref is not there.  Returns dict(value GB/s, cores, kind, sample, ...)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    words_per_file = chunks_per_file * CHUNK_WORDS
    if O.have_ref():
        tmp = tempfile.mkdtemp(prefix="mrcz_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            src_dir, zip_dir, out_dir = Path(tmp) / "src", Path(tmp) / "zip", Path(tmp) / "out"
            for d in (src_dir, zip_dir, out_dir):
                d.mkdir()
            names = []
            for i in range(cores):
                w = sample_words_fn(words_per_file, i)
                p = src_dir / f"v{i:03d}.mrc"
                w.tofile(p)
                names.append(p)
            total = cores * words_per_file * 4
            (Path(tmp) / "zip.txt").write_text("".join(f"{p}\n" for p in names))
            (Path(tmp) / "unzip.txt").write_text("".join(f"{zip_dir / (p.name + '.zip')}\n" for p in names))
            exe = str(O.REF_DIR / "mrc_tarx_c")

            def one():
                t0 = time.perf_counter()
                subprocess.run([exe, "-i", str(Path(tmp) / "zip.txt"), "-t", "zip", "-o", str(zip_dir), "-b", str(bits),
                                "-n", str(cores), "-d", "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                t1 = time.perf_counter()
                subprocess.run([exe, "-i", str(Path(tmp) / "unzip.txt"), "-t", "unzip", "-o", str(out_dir),
                                "-n", str(cores), "-d", "1"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                t2 = time.perf_counter()
                return t1 - t0, t2 - t1

            for _ in range(warmup):
                one()
            tz, tu = [], []
            for _ in range(steps):
                a, b = one()
                tz.append(a); tu.append(b)
            zsize = sum((zip_dir / (p.name + ".zip")).stat().st_size - 17 for p in names)
            tzm, tum = float(np.mean(tz)), float(np.mean(tu))
            return dict(value=total / (tzm + tum) / 1e9, unit=UNIT, cores=cores, kind="reference",
                        sample=f"{cores} files x {chunks_per_file} chunks ({total / 2**30:.2f} GiB {kind} b={bits}) in /dev/shm, "
                               f"mrc_tarx_c -n {cores} zip (-d 0) + unzip (-d 1), reference built -O2 with its own zlib 1.2.8",
                        compress_GBs=total / tzm / 1e9, decompress_GBs=total / tum / 1e9, ratio=zsize / total,
                        ms_per_step=(tzm + tum) * 1e3, bytes=total)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    # port: the oracle restatement, one thread
    w = sample_words_fn(min(words_per_file, 3 * CHUNK_WORDS), 0)
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


def host_sample_fn(kind: str):
    """Sample generator for the CPU legs (numpy / torch-CPU, no GPU needed)."""
    import torch

    def fn(nwords: int, i: int) -> np.ndarray:
        g = torch.Generator()
        g.manual_seed({"G": 1234, "P": 4321, "S": 7}[kind] + 1000 + i)
        if kind == "G":
            d = torch.randn(nwords, generator=g, dtype=torch.float32)
        elif kind == "P":
            d = torch.poisson(torch.full((nwords,), 2.0), generator=g)
        else:
            x = torch.linspace(0, 4000 * np.pi, nwords)
            d = torch.sin(x) * torch.cos(x / 7) + 0.25 * torch.randn(nwords, generator=g)
        w = d.numpy().view(np.uint32).copy()
        w[:256] = 0
        return w
    return fn


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference_run(a.kind, a.bits, a.cpu_chunks_per_file, a.steps, a.warmup, host_sample_fn(a.kind))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(a, gpu_words_per_rank(a, 0)),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "compress_GBs": r.get("compress_GBs"), "decompress_GBs": r.get("decompress_GBs"), "ratio": r.get("ratio"),
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def gpu_words_per_rank(a, rank):
    data_words = int(a.gib * (1 << 30)) // 4
    data_words = (data_words // CHUNK_WORDS) * CHUNK_WORDS if data_words >= CHUNK_WORDS else data_words
    return data_words + (256 if rank == 0 else 0)


def workload_config(a, nwords_per_gpu):
    return {"workload": f"mrc_full-style round trip, synthetic {a.gib:g} GiB float32 shard per GPU "
                        f"(1024^3 MRC volume at 4 GiB; {a.gpus} x {a.gib:g} GiB volume sharded by contiguous chunk range)",
            "distribution": {"G": "normal(0,1)", "P": "poisson(2)", "S": "smooth+noise"}[a.kind], "mask_bits": a.bits,
            "chunk_words": CHUNK_WORDS, "words_per_gpu": int(nwords_per_gpu), "zlib_equivalent": "level 6, Z_RLE, raw deflate",
            "l2_policy": "inputs (>= 1 GiB per launch) exceed the 126 MB L2; no flush needed", "parallelism": f"chunk-range x{a.gpus}"}


def bind_near_gpu(local_rank):
    """Multi-rank runs: keep this process (and the pinned host buffers it is about to allocate) on the NUMA node
    of its GPU, so that eight ranks do not pull 100+ GB per step across the socket interconnect."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            return {"numa_node": None}
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # no sysfs / unknown properties: run unbound
        return {"numa_node": None, "note": type(e).__name__}


# ----------------------------------------------------------------------------- the B200 arm
def run_b200(a):
    import torch
    import torch.distributed as dist
    from datacompressionfloat_b200 import Codec

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

    # rank 0 carries the 1024-byte MRC header in front of its chunk range (header words are exempt from masking)
    nwords = gpu_words_per_rank(a, rank)
    data_words = nwords - (256 if rank == 0 else 0)
    words = torch.empty(nwords, dtype=torch.int32, device=dev)
    words[nwords - data_words:] = gen_words(a.kind, data_words, rank, dev)
    if rank == 0:
        words[:256] = 0
        words[0:3] = 1024
        words[3] = 2
    exempt = 256 if rank == 0 else 0
    codec = Codec.on_current_stream(batch_chunks=a.batch_chunks or None)
    codec.set_profiling(True)
    cont_buf = torch.empty(Codec.compress_bound(nwords), dtype=torch.uint8, device=dev)
    out_words = torch.empty(nwords, dtype=torch.int32, device=dev)
    sizes_all = torch.zeros(world, dtype=torch.int64, device=dev)
    my_size = torch.zeros(1, dtype=torch.int64, device=dev)
    state = {}

    def step(timed=None):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed is not None else None
        if e: e[0].record()
        seg = codec.compress(words, a.bits, exempt_words=exempt, write_file_header=False, out=cont_buf)
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

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # samples clocks / throttle reasons from the warm-up through the timed region
    for _ in range(a.warmup):
        step()
    # correctness of what is being timed: masked round trip, bit exact
    mask = (-1 << a.bits) if a.bits < 32 else 0
    ref = words.clone()
    ref[exempt:] &= mask
    ok = bool(torch.equal(ref, state["back"]))
    del ref
    if not ok:
        raise SystemExit("round trip is not bit-exact: refusing to report a number")

    timed = []
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    t0.record()
    for _ in range(a.steps):
        step(timed)
    t1.record()
    barrier()
    sampler.timed = (wall0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
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
    ms_per_step = elapsed_ms / a.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with HOST buffers (pinned), H2D and D2H inside the timed region
    e2e = None
    if not a.no_e2e:
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
        del h_in, h_cont, h_out

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
    ratio = state["seg"].numel() / (nwords * 4.0)
    nb = nwords * 4.0
    alg = {  # algorithmic bytes per launch group, r = compressed / original
        "split": 2 * nb, "merge": 2 * nb, "encode": nb + nb * ratio, "gather": 2 * nb * ratio,
        "inflate_fast": nb * ratio + nb, "rawcopy": 2 * nb * ratio, "markers": nb * ratio,
    }
    stages = {k: stage_c.get(k, 0.0) + stage_d.get(k, 0.0) for k in set(stage_c) | set(stage_d)}
    kernels = []
    for k, ms in sorted(stages.items(), key=lambda kv: -kv[1]):
        if ms <= 0 or k not in alg:
            continue
        ach = alg[k] / (ms * 1e-3) / 1e9
        kernels.append({"kernel": k, "ms_per_step": ms, "achieved": ach, "frac": ach / peak, "algorithmic_bytes": alg[k]})
    dom = kernels[0] if kernels else None
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and dom:
        try:
            traffic = json.loads(tf.read_text()).get(dom["kernel"])
        except Exception:
            traffic = None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac"], "traffic": traffic, "peak_source": peak_src,
                    "note": "achieved = algorithmic bytes of the stage per step / its CUDA-event time (all launches of the stage in a step)"}

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        try:
            r = cpu_reference_run(a.kind, a.bits, a.cpu_chunks_per_file, 1, 0, host_sample_fn(a.kind))
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu.update(compress_GBs=r.get("compress_GBs"), decompress_GBs=r.get("decompress_GBs"), ratio=r.get("ratio"))
        except Exception as ex:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    # whole-pipeline view: bytes a perfectly fused implementation would have to move (SURVEY 8d lower bound,
    # 4 + 4r per word each way) over the measured step time, as a fraction of the measured HBM peak
    fused_bytes = 2 * (nb + nb * ratio) * world
    pipeline = {"fused_lower_bound_bytes": fused_bytes, "achieved": fused_bytes / (ms_per_step * 1e-3) / 1e9 / world,
                "unit": "GB/s per GPU", "frac_of_hbm_peak": fused_bytes / (ms_per_step * 1e-3) / 1e9 / world / peak}
    launches = (state["cs"]["kernel_launches"] + state["ds"]["kernel_launches"]) * a.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(a, nwords),
        "compress_GBs": total_bytes / (comp_ms * 1e-3) / 1e9, "decompress_GBs": total_bytes / (decomp_ms * 1e-3) / 1e9,
        "ratio": total_comp / total_bytes, "ratio_definition": "compressed/original (reference zip.c:434), chunk records only",
        "bit_exact_roundtrip": ok, "roofline": roofline, "roofline_kernels": kernels, "pipeline_roofline": pipeline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
        "decode_stats": {k: state["ds"][k] for k in ("general_streams", "fast_failed")},
        "host_affinity": affinity,
        "encode_stats": {k: state["cs"][k] for k in ("raw_streams", "stored_subblocks", "streams")},
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
