"""Minimal driver for profiling: sets up one 8K (or other) frame and runs N resident encode+decode steps.
    ncu --set full --clock-control none --import-source on -k regex:k_ -s 12 -c 6 -o gpurun_out/prof python profiles/run_step.py
Kernel launch order per step: k_fdct_rgb444, k_huff_encode_packed, k_huff_place, k_marker_scan_write,
k_huff_decode_sync, k_idct_rgb444 (6 launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import _oracle as o  # noqa: E402  (synthetic frame generator only)
import gpujpeg_b200 as g  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "8k"
kind = sys.argv[2] if len(sys.argv) > 2 else "photo"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
w, h, rst = {"8k": (7680, 4320, 36), "4k": (3840, 2160, 24), "hd": (1920, 1080, 24), "16k": (15360, 8640, 36)}[size]
img = o.gen_image(kind, w, h)
d_raw = torch.from_numpy(img).cuda()
enc, dec = g.Encoder(), g.Decoder()
jpeg = enc.encode(d_raw, 75, rst)
d_out = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
dec.decode(jpeg, out=d_out)
torch.cuda.synchronize()
for _ in range(steps):
    enc.run_resident(d_raw, 3)
    dec.run_resident(d_out, 7)
torch.cuda.synchronize()
print("done", jpeg.size)
