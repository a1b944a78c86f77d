"""Generates tests/golden/*.npz from the REFERENCE's own CPU code (oracle/_ref/libgpujpeg_refcpu.so,
compiled in place from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds: the synthetic RGB input recipe, the quantised coefficients fed to the reference
Huffman encoder, the JPEG bytes the reference header writer + CPU Huffman encoder produced, the
coefficients the reference CPU Huffman decoder recovered, and the planes the reference integer IDCT
(gpujpeg_idct_cpu_perform, +128, clamp) produced.  The forward DCT / colour stages have no CPU
implementation in the reference, so the coefficients fed in here come from the oracle restatement; that
restatement is pinned separately against committed outputs of the reference GPU library
(tests/golden/refgpu_*.npz, made by make_golden_refgpu.py on a B200) and live on the GPU box (tests/test_ref_gpu.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as o  # noqa: E402
from _refcpu import ref_decode_coef, ref_encode_coef, ref_idct_planes  # noqa: E402

CASES = [  # name, kind, w, h, quality, rst, interleaved
    ("random_64x48_q75_r4", "random", 64, 48, 75, 4, 0),
    ("photo_96x64_q75_r24", "photo", 96, 64, 75, 24, 0),
    ("random_33x17_q90_r2", "random", 33, 17, 90, 2, 0),
    ("gradient_80x40_q50_r8", "gradient", 80, 40, 50, 8, 0),
    ("random_40x24_q100_r1", "random", 40, 24, 100, 1, 0),
    ("photo_72x56_q75_r5_il", "photo", 72, 56, 75, 5, 1),
    ("random_48x32_q20_r0", "random", 48, 32, 20, 0, 0),
]


SS_CASES = [  # name, kind, w, h, quality, rst, interleaved, luminance sampling (chrominance 1x1)
    ("ss420_photo_72x56_q75_r3_il", "photo", 72, 56, 75, 3, 1, (2, 2)),
    ("ss422_random_33x17_q90_r2", "random", 33, 17, 90, 2, 0, (2, 1)),
    ("ss440_random_40x24_q60_r0_il", "random", 40, 24, 60, 0, 1, (1, 2)),
    ("ss420_random_50x30_q85_r4", "random", 50, 30, 85, 4, 0, (2, 2)),
]


def main():
    assert o.ref is not None, "build oracle/_ref first (make -C oracle ref)"
    for name, kind, w, h, q, rst, il, samp in SS_CASES:
        img = o.gen_image(kind, w, h)
        _, coef = o.encode(img, q, rst, il, want_coef=True, sampling=samp)
        jpeg = ref_encode_coef(coef, w, h, q, rst, il, sampling=samp)
        coef_dec = ref_decode_coef(jpeg, w, h, rst, il, sampling=samp)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=kind, w=w, h=h, quality=q, rst=rst, interleaved=il,
                            sampling=np.array(samp), coef=coef, jpeg=jpeg, coef_dec=coef_dec)
        print(name, "jpeg bytes", jpeg.size)
    for name, kind, w, h, q, rst, il in CASES:
        img = o.gen_image(kind, w, h)
        _, coef = o.encode(img, q, rst, il, want_coef=True)
        jpeg = ref_encode_coef(coef, w, h, q, rst, il)
        coef_dec = ref_decode_coef(jpeg, w, h, rst, il)
        planes = ref_idct_planes(coef_dec, w, h, q)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=kind, w=w, h=h, quality=q, rst=rst,
                            interleaved=il, coef=coef, jpeg=jpeg, coef_dec=coef_dec, planes=planes)
        print(name, "jpeg bytes", jpeg.size)


if __name__ == "__main__":
    main()
