"""Frame-level sharding of a batch over the GPUs of one box.

Frames are independent, so the codec's data path has NO collective: rank r simply encodes/decodes the frames
it owns with its own coder instance.  Communication exists only when the frames originate on (or the results
are wanted on) one rank: a scatter of raw frames and a gather-v of the variable-length JPEG byte strings,
both expressed with point-to-point `torch.distributed` ops so they run over NCCL/NVLink on GPUs and over
gloo on CPUs (tests/test_batch_gloo.py uses world_size 2 with the CPU oracle standing in for the coder).
"""
import torch
import torch.distributed as dist


def owner(frame, world):
    """round-robin: frame f lives on rank f % world"""
    return frame % world


def my_frames(n_frames, world, rank):
    return list(range(rank, n_frames, world))


def scatter_frames(frames, n_frames, shape, src=0, device=None, group=None):
    """frames: list of n_frames uint8 tensors of `shape` on rank `src` (ignored elsewhere).
    Returns the list of tensors this rank owns (in frame order)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = my_frames(n_frames, world, rank)
    if rank == src:
        ops, out = [], []
        for f in range(n_frames):
            o = owner(f, world)
            if o == src:
                out.append(frames[f] if device is None else frames[f].to(device))
            else:
                ops.append(dist.P2POp(dist.isend, frames[f].contiguous(), o, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out
    out = [torch.empty(shape, dtype=torch.uint8, device=device) for _ in mine]
    ops = [dist.P2POp(dist.irecv, t, src, group) for t in out]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def gather_streams(streams, n_frames, dst=0, device=None, group=None):
    """streams: this rank's JPEG byte strings (1-D uint8 tensors) in frame order.
    Returns on `dst` the list of all n_frames streams in frame order, elsewhere None."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_frames + world - 1) // world
    sizes = torch.zeros(per, dtype=torch.int64, device=device)
    for i, s in enumerate(streams):
        sizes[i] = s.numel()
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    if rank == dst:
        result = [None] * n_frames
        ops, bufs = [], {}
        for r in range(world):
            fr = my_frames(n_frames, world, r)
            if r == dst:
                for i, f in enumerate(fr):
                    result[f] = streams[i]
                continue
            total = int(all_sizes[r][:len(fr)].sum().item())
            if total == 0:
                continue
            bufs[r] = torch.empty(total, dtype=torch.uint8, device=device)
            ops.append(dist.P2POp(dist.irecv, bufs[r], r, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for r, buf in bufs.items():
            off = 0
            for i, f in enumerate(my_frames(n_frames, world, r)):
                n = int(all_sizes[r][i].item())
                result[f] = buf[off:off + n]
                off += n
        return result
    if streams:
        payload = torch.cat([s.reshape(-1) for s in streams]) if len(streams) > 1 else streams[0].reshape(-1).contiguous()
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, payload, dst, group)]):
            w.wait()
    return None


class BatchCodec:
    """One coder pair per process (= per GPU).  encode()/decode() work on this rank's frames only."""

    def __init__(self, stream=0):
        from . import api
        self.enc = api.Encoder(stream=stream, pinned_output=True)
        self.dec = api.Decoder(stream=stream)

    def encode(self, frames, quality=75, restart_interval=-1):
        """frames: list of HxWx3 uint8 tensors/arrays (host or cuda) -> list of numpy uint8 JPEG streams"""
        return [self.enc.encode(f, quality, restart_interval) for f in frames]

    def decode(self, streams, outs=None):
        if outs is None:
            return [self.dec.decode(s) for s in streams]
        return [self.dec.decode(s, out=o) for s, o in zip(streams, outs)]

    def close(self):
        self.enc.close()
        self.dec.close()
