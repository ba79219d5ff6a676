"""GPU aid: frame-sharded detection over N ranks (one process per GPU, no data-path collective) gives the same per-frame
results as the oracle and as one GPU running the whole batch.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/shard_parity_gpu.py [frames]
Every rank detects its contiguous slice (rmcv_b200.shard.frame_slice) on its own GPU; per-frame digests (mask CRC, counts,
raw bytes of the blob and armour records) are gathered on the host in rank order; rank 0 also runs the whole batch on
its GPU and the oracle on every frame."""
import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import rmcv_b200 as rb
from rmcv_b200 import shard, synth

total = int(sys.argv[1]) if len(sys.argv) > 1 else 67      # not a multiple of the rank count on purpose
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
seeds = list(range(5000, 5000 + total))


def digests(device, lo, hi):
    frames = np.stack([synth.make_frame(s, 1280, 1024, synth.plates_for_seed(s)) for s in seeds[lo:hi]])
    out = []
    with rb.Context(max_width=1280, max_height=1024, max_batch=max(1, hi - lo), device=device, chunk_frames=5) as c:
        masks = np.empty((hi - lo, 1024, 1280), np.uint8)
        res = c.detect_batch_host(frames, rb.default_params(), masks)
        for f in range(hi - lo):
            d = c.frame_detections(res, f)
            blob = b"".join(np.asarray([b.angle, *b.center, *np.asarray(b.vertices).ravel(), *b.size], np.float32).tobytes() for b in d.positive)
            arm = b"".join(np.asarray(a.vertices, np.float32).tobytes() + np.asarray(a.icon, np.float32).tobytes() for a in d.armours)
            out.append((zlib.crc32(masks[f].tobytes()), len(d.contours), len(d.positive), len(d.armours), d.n_negative,
                        [tuple(ci.first) + (ci.n_points, ci.area2) for ci in d.contours], zlib.crc32(blob), zlib.crc32(arm)))
    return out


lo, hi = shard.frame_slice(total, world, rank)
mine = digests(local, lo, hi)
if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
else:
    gathered = [mine]
ok = True
if rank == 0:
    sharded = [d for part in gathered for d in part]
    whole = digests(local, 0, total)
    ok = sharded == whole and len(sharded) == total
    from oracle import rm_oracle as O
    for k, s in enumerate(seeds):
        ref = O.detect_frame(synth.make_frame(s, 1280, 1024, synth.plates_for_seed(s)))
        ok = ok and sharded[k][0] == zlib.crc32(ref.binary.tobytes()) and sharded[k][1:4] == (len(ref.contours), len(ref.positive), len(ref.armours))
    print("shard_parity: %d frames over %d ranks (slices %s): sharded == one GPU == oracle: %s" % (total, world, shard.all_slices(total, world), ok))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
