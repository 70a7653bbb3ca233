"""BASELINE.json configs[4] in miniature, with the full cross-decode (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      tools/config5_crosscheck.py --gib-per-gpu 1

Every rank compresses its contiguous chunk range; the per-rank segment sizes are all_gathered (the only
exchange), their exclusive scan gives the container offsets, every rank writes its segment at its offset
into ONE container file; rank 0 then lets the REFERENCE binary (mrc_tar_c -t unzip, its own libz 1.2.8)
inflate that file and compares with the reference's erasebytes_c output of the original volume.
"""
import argparse, json, os, subprocess, sys, time
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from datacompressionfloat_b200 import Codec, CHUNK_WORDS, chunk_range, segment_offsets, file_header  # noqa: E402
from oracle import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gib-per-gpu", type=float, default=1.0)
ap.add_argument("--bits", type=int, default=8)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
chunks_per_rank = max(1, int(a.gib_per_gpu * (1 << 30)) // 4 // CHUNK_WORDS)
nchunks = chunks_per_rank * world
total_words = nchunks * CHUNK_WORDS - 12345          # ragged last chunk
lo, hi = chunk_range(nchunks, rank, world)
w0, w1 = lo * CHUNK_WORDS, min(total_words, hi * CHUNK_WORDS)
g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
words = torch.randn(w1 - w0, generator=g, device="cuda").view(torch.int32)
if rank == 0:
    words[:256] = 7                                    # MRC header words: must survive unmasked
tmp = "/dev/shm/cfg5"
if rank == 0:
    os.makedirs(tmp, exist_ok=True)
dist.barrier()
vol, cont = f"{tmp}/volume.mrc", f"{tmp}/volume.mrc.zip"
codec = Codec.on_current_stream()
seg = codec.compress(words, a.bits, exempt_words=256 if rank == 0 else 0, write_file_header=False)
sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
dist.all_gather_into_tensor(sizes, torch.tensor([seg.numel()], dtype=torch.int64, device="cuda"))
offs, total = segment_offsets(sizes.tolist())
if rank == 0:
    with open(cont, "wb") as f:
        f.truncate(total)
    with open(vol, "wb") as f:
        f.truncate(total_words * 4)
dist.barrier()
mm = np.memmap(cont, dtype=np.uint8, mode="r+")
if rank == 0:
    mm[:17] = file_header(total_words * 4)
mm[offs[rank]: offs[rank] + seg.numel()] = seg.cpu().numpy()
mm.flush(); del mm
mv = np.memmap(vol, dtype=np.int32, mode="r+")
mv[w0:w1] = words.cpu().numpy()
mv.flush(); del mv
# shard-wise decode on the GPUs as well
back = codec.decompress(seg, has_file_header=False, nwords=w1 - w0)
ref = words.clone(); ref[256 if rank == 0 else 0:] &= (-1 << a.bits)
ok_local = torch.tensor([int(torch.equal(ref, back))], device="cuda")
dist.all_reduce(ok_local, op=dist.ReduceOp.MIN)
dist.barrier()
if rank == 0:
    rec = dict(config=5, commit=os.environ.get("MRCZIP_COMMIT"), n_gpus=world, volume_bytes=total_words * 4, container_bytes=total, ratio=(total - 17) / (total_words * 4),
               gpu_shard_roundtrip_ok=bool(ok_local.item()))
    if O.have_ref():
        t0 = time.time()
        subprocess.run([O.REF_DIR / "mrc_tar_c", "-i", cont, "-o", f"{tmp}/ref.out", "-t", "unzip"], check=True, stdout=subprocess.DEVNULL)
        t_ref = time.time() - t0
        subprocess.run([O.REF_DIR / "erasebytes_c", "-i", vol, "-o", f"{tmp}/golden.mrc", "-b", str(a.bits)], check=True, stdout=subprocess.DEVNULL)
        same = subprocess.run(["cmp", "-s", f"{tmp}/ref.out", f"{tmp}/golden.mrc"]).returncode == 0
        rec.update(reference_unzip_of_gpu_container_equals_erasebytes=same, reference_unzip_s=t_ref)
    print(json.dumps(rec), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/config5.json").write_text(json.dumps(rec))
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    subprocess.run(["rm", "-rf", tmp])
dist.destroy_process_group()
