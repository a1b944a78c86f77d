"""Multi-process sharding logic (gpujpeg_b200/batch.py) on CPU: world_size 2, gloo backend, the CPU oracle
standing in for the per-rank coder.  Checks that scatter -> per-rank encode -> gather-v returns every
frame's stream, in order, identical to encoding the batch sequentially.  CPU only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, w, h, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("gj_batch", os.path.join(ROOT, "gpujpeg_b200", "batch.py"))
    batch = importlib.util.module_from_spec(spec)   # import batch.py alone: no GPU library needed for the transport
    spec.loader.exec_module(batch)
    import _oracle as o
    frames = None
    if rank == 0:
        frames = [torch.from_numpy(o.gen_image("photo" if f % 2 else "random", w, h, seed=100 + f)) for f in range(n_frames)]
    mine = batch.scatter_frames(frames, n_frames, (h, w, 3), src=0)
    assert len(mine) == len(batch.my_frames(n_frames, world, rank))
    streams = [torch.from_numpy(o.encode(f.numpy(), q, 4)) for f in mine]
    got = batch.gather_streams(streams, n_frames, dst=0)
    if rank == 0:
        for f in range(n_frames):
            want = o.encode(frames[f].numpy(), q, 4)
            assert np.array_equal(got[f].numpy(), want), "frame %d" % f
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [1, 4, 5])
def test_scatter_encode_gather_world2(n_frames):
    mp.spawn(_worker, args=(2, _free_port(), n_frames, 96, 64, 75), nprocs=2, join=True)


def test_shard_assignment():
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("gj_batch", os.path.join(ROOT, "gpujpeg_b200", "batch.py"))
    batch = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(batch)
    for world in (1, 2, 4, 8):
        seen = sorted(f for r in range(world) for f in batch.my_frames(32, world, r))
        assert seen == list(range(32))
        assert all(batch.owner(f, world) == r for r in range(world) for f in batch.my_frames(32, world, r))
