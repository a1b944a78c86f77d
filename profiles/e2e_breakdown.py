"""Prints the per-stage breakdown of the reference-facing API calls (host buffers) plus raw PCIe copy rates.
    python profiles/e2e_breakdown.py [8k] [photo]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import _oracle as o  # noqa: E402
import gpujpeg_b200 as g  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "8k"
kind = sys.argv[2] if len(sys.argv) > 2 else "photo"
w, h, rst = {"8k": (7680, 4320, 36), "4k": (3840, 2160, 24), "hd": (1920, 1080, 24), "16k": (15360, 8640, 36)}[size]
img = torch.from_numpy(o.gen_image(kind, w, h)).pin_memory()
dev = torch.empty_like(img, device="cuda")
for name, fn in (("H2D pinned", lambda: dev.copy_(img, non_blocking=True)), ("D2H pinned", lambda: img.copy_(dev, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 10
    print("%s: %.3f ms for %.1f MB = %.1f GB/s" % (name, dt * 1e3, img.numel() / 1e6, img.numel() / dt / 1e9))
enc, dec = g.Encoder(pinned_output=True), g.Decoder()
p = g.api.default_parameters(75, rst)
p.perf_stats = 1
pi = g.api.image_parameters(w, h)
h_out = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
host = img.numpy()
for it in range(4):
    t0 = time.perf_counter()
    addr, size_ = enc.encode_raw(host, p, pi, device=False)
    t1 = time.perf_counter()
    dec.decode_raw(addr, size_, g.api.GPUJPEG_DECODER_OUTPUT_CUSTOM_BUFFER, h_out.data_ptr())
    t2 = time.perf_counter()
    se = enc.stats()
    print("iter %d: encode %.3f ms  decode %.3f ms  jpeg %d B" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, size_))
    if se:
        print("   enc: to %.3f  dct %.3f  huff %.3f  from %.3f  stream %.3f  in_gpu %.3f" %
              (se.duration_memory_to, se.duration_dct_quantization, se.duration_huffman_coder, se.duration_memory_from,
               se.duration_stream, se.duration_in_gpu))
# decoder stats need perf_stats on the decoder: use verbose status print once
import ctypes as C
d2 = g.lib.gpujpeg_decoder_create(None)
pp = g.api.default_parameters(75, rst); pp.verbose = 1; pp.comp_count = 3
g.lib.gpujpeg_decoder_init.argtypes = [C.c_void_p, C.POINTER(g.api.Parameters), C.POINTER(g.api.ImageParameters)]
g.lib.gpujpeg_decoder_init(d2, C.byref(pp), C.byref(pi))
out = g.api.DecoderOutput(); out.type = g.api.GPUJPEG_DECODER_OUTPUT_CUSTOM_BUFFER; out.data = h_out.data_ptr()
for _ in range(3):
    g.lib.gpujpeg_decoder_decode(d2, C.c_void_p(addr), size_, C.byref(out))
