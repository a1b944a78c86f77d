"""K2 alone (encode + place) and its tail: python profiles/k2_times.py [size] [kind]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _oracle as o
import gpujpeg_b200 as g
size = sys.argv[1] if len(sys.argv) > 1 else "8k"
kind = sys.argv[2] if len(sys.argv) > 2 else "photo"
w, h, rst = {"8k": (7680, 4320, 36), "4k": (3840, 2160, 24), "hd": (1920, 1080, 24)}[size]
d_raw = torch.from_numpy(o.gen_image(kind, w, h)).cuda()
enc = g.Encoder(stream=torch.cuda.current_stream().cuda_stream)
ref = enc.encode(d_raw, 75, rst, 0)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(size, kind, "K1 %.1f us  K2 %.1f us  K1+K2 %.1f us  (jpeg %d B)" % (timeit(lambda: enc.run_resident(d_raw, 1)), timeit(lambda: enc.run_resident(d_raw, 2)),
                                                                     timeit(lambda: enc.run_resident(d_raw, 3)), ref.size))
import numpy as np
assert np.array_equal(enc.encode(d_raw, 75, rst, 0), ref)
