"""gpujpegx_batch_*: a batch of independent frames sharded over workers (one per device entry; a device may be listed
more than once, so the sharding, the threading and the peer-copy paths are exercised on a single-GPU box too).
Every frame's stream and pixels must equal what the oracle gives for that frame alone.  GPU, through the C ABI."""
import os
import subprocess

import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def frames(n, w, h):
    return [o.gen_image("photo" if f % 2 else "random", w, h, seed=500 + f) for f in range(n)]


@pytest.mark.parametrize("workers,n", [(1, 3), (2, 5), (3, 7)])
def test_batch_host_frames(workers, n):
    import gpujpeg_b200 as g
    w, h = 320, 200
    imgs = frames(n, w, h)
    b = g.api.Batch([0] * workers)
    try:
        assert b.device_count == workers and b.owner(1) == 0
        p = g.api.default_parameters(80, 6)
        jpegs = b.encode(imgs, p, g.api.image_parameters(w, h))
        for f in range(n):
            assert np.array_equal(jpegs[f], o.encode(imgs[f], 80, 6)), "frame %d" % f
        outs = [np.empty((h, w, 3), np.uint8) for _ in range(n)]
        b.decode(jpegs, outs)
        for f in range(n):
            assert np.array_equal(outs[f], o.decode(jpegs[f])), "frame %d" % f
        assert b.last_ms() > 0
        # a second batch with other parameters on the same workers (coder re-initialisation inside the batch)
        imgs2 = frames(4, 96, 64)
        j2 = b.encode(imgs2, g.api.default_parameters(60, 2, 1, "4:2:0"), g.api.image_parameters(96, 64))
        for f in range(4):
            assert np.array_equal(j2[f], o.encode(imgs2[f], 60, 2, 1, sampling=(2, 2)))
    finally:
        b.close()


@pytest.mark.parametrize("where", ["owner", "first"])
def test_batch_device_frames(where):
    """frames resident on the owning device, or all on the first device and moved over the peer path"""
    import torch

    import gpujpeg_b200 as g
    w, h, n = 256, 128, 6
    imgs = frames(n, w, h)
    loc = g.api.GPUJPEGX_DEVICE_OWNER if where == "owner" else g.api.GPUJPEGX_DEVICE_FIRST
    b = g.api.Batch([0, 0])
    try:
        d_imgs = [torch.from_numpy(i).cuda() for i in imgs]
        jpegs = b.encode(d_imgs, g.api.default_parameters(75, 8), g.api.image_parameters(w, h), where=loc)
        for f in range(n):
            assert np.array_equal(jpegs[f], o.encode(imgs[f], 75, 8))
        d_out = [torch.empty((h, w, 3), dtype=torch.uint8, device="cuda") for _ in range(n)]
        b.decode(jpegs, d_out, where=loc)
        torch.cuda.synchronize()
        for f in range(n):
            assert np.array_equal(d_out[f].cpu().numpy(), o.decode(jpegs[f]))
    finally:
        b.close()


def test_batch_from_plain_c(tmp_path):
    from gpujpeg_b200 import build as bld
    lib = bld.build_library()
    exe = str(tmp_path / "batch")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", os.path.join(HERE, "c_api", "batch.c"), "-I",
                           os.path.join(ROOT, "include"), "-L", os.path.dirname(lib), "-lgpujpeg", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(lib) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    w, h, n = 200, 120, 5
    out = subprocess.check_output([exe, str(w), str(h), str(n), str(tmp_path), "0", "0"], env=env).decode()
    assert out.startswith("5 frames on 2 workers")
    for f in range(n):
        y, x, c = np.meshgrid(np.arange(h), np.arange(w), np.arange(3), indexing="ij")
        img = ((y * 255 // h + 17 * f + 40 * c + (x & 7)) & 255).astype(np.uint8)
        want = o.encode(np.ascontiguousarray(img), 80, 4)
        assert np.array_equal(np.fromfile(tmp_path / ("%d.jpg" % f), np.uint8), want)
        assert np.array_equal(np.fromfile(tmp_path / ("%d.rgb" % f), np.uint8).reshape(h, w, 3), o.decode(want))
