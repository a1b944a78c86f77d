"""SURVEY 8e, "frames originate on GPU 0": NCCL scatter of raw frames from rank 0, encode on every rank, gather-v of the
JPEG byte strings back to rank 0 -- next to the same work with the frames already resident on their GPUs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/scatter_bench.py [frames_per_rank]
Rank 0 checks every gathered stream against its own encoding of the same frame and prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import _oracle as o  # noqa: E402  (frame generator only)
import gpujpeg_b200 as g  # noqa: E402
from gpujpeg_b200 import batch  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w, h, rst = 7680, 4320, 36
n = per_rank * world
codec = batch.BatchCodec(stream=torch.cuda.current_stream().cuda_stream)
frames = None
if rank == 0:
    base = torch.from_numpy(o.gen_image("photo", w, h)).to(dev)
    frames = [torch.roll(base, shifts=37 * f, dims=1) for f in range(n)]   # n different frames, all on GPU 0
    expect = [codec.enc.encode(f, 75, rst) for f in frames]


def sync():
    dist.barrier()
    torch.cuda.synchronize()


def run(scatter):
    sync()
    t0 = time.perf_counter()
    mine = batch.scatter_frames(frames, n, (h, w, 3), 0, dev) if scatter else resident
    streams = [torch.from_numpy(s).to(dev) for s in codec.encode(mine, 75, rst)]
    out = batch.gather_streams(streams, n, 0, dev)
    sync()
    return time.perf_counter() - t0, mine, out


_, resident, _ = run(True)      # warm-up; also leaves this rank's frames resident for the second measurement
t_scatter, _, got = run(True)
t_resident, _, got2 = run(False)
if rank == 0:
    for f in range(n):
        assert np.array_equal(got[f].cpu().numpy(), expect[f]), "frame %d: gathered stream differs" % f
        assert np.array_equal(got2[f].cpu().numpy(), expect[f])
    mpix = n * w * h / 1e6
    print(json.dumps({"n_gpus": world, "frames": n, "frame": "%dx%d" % (w, h),
                      "encode_scattered_from_gpu0_mpix_s": round(mpix / t_scatter, 1),
                      "encode_resident_mpix_s": round(mpix / t_resident, 1),
                      "scatter_ms": round((t_scatter - t_resident) * 1e3, 3), "checked": "all gathered streams == rank 0's own"}))
codec.close()
dist.destroy_process_group()
