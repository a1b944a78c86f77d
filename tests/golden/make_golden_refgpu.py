"""Generates tests/golden/refgpu_*.npz ON A GPU BOX: outputs of the REFERENCE GPU library itself (oracle/_ref/
libgpujpeg_refgpu.so = the reference's own sources compiled in place by `make -C oracle refgpu`) for a handful of small
seeded inputs.  The colour transform and the float forward DCT exist in the reference only as CUDA kernels, so this is
the one place their results can be taken from; the committed fixtures then pin oracle.c's restatement of that arithmetic
in the GPU-less test run (tests/test_oracle_golden.py) -- no "the golden coefficients come from the oracle" caveat.

    python tests/golden/make_golden_refgpu.py [outdir [name-suffix]]   (default: gpurun_out/golden, which gpurun brings back;
                                                                        with a suffix only the cases whose name ends in it)

Each fixture: the generator call that makes the input (kind, w, h -> _oracle.gen_image), the encoder parameters, the
reference GPU encoder's JPEG bytes, and the reference GPU decoder's pixels for that stream (its float IDCT)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)

CASES = [  # name, kind, w, h, quality, rst, interleaved, luminance sampling (h, v)
    ("random_1119x561_q75", "random", 1119, 561, 75, 8, 0, (1, 1)),     # the reference's own regression input
    ("photo_640x360_q90_il", "photo", 640, 360, 90, 7, 1, (1, 1)),
    ("random_200x120_q100", "random", 200, 120, 100, 36, 0, (1, 1)),    # long codes, stuffing
    ("photo_322x200_420_il", "photo", 322, 200, 80, 6, 1, (2, 2)),
    ("random_161x97_422", "random", 161, 97, 85, 4, 0, (2, 1)),
    ("photo_128x64_440_il", "photo", 128, 64, 75, 2, 1, (1, 2)),
    ("gradient_640x480_q75", "gradient", 640, 480, 75, 8, 0, (1, 1)),
    ("random_33x17_q90", "random", 33, 17, 90, 2, 0, (1, 1)),
    ("photo_64x48_q1", "photo", 64, 48, 1, 5, 0, (1, 1)),
    ("zero_256x256", "zero", 256, 256, 75, 36, 0, (1, 1)),
    # param.segment_info = 1: APP13 tables of the restart segments' positions in front of every scan
    ("photo_256x192_seginfo", "photo", 256, 192, 75, 4, 0, (1, 1), 1),
    ("photo_320x208_420_seginfo", "photo", 320, 208, 75, 2, 0, (2, 2), 1),
    ("random_200x120_il_seginfo", "random", 200, 120, 90, 3, 1, (1, 1), 1),
]


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(TESTS), "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    srv = subprocess.Popen([sys.executable, os.path.join(TESTS, "_refgpu.py"), "serve"], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                           text=True)
    assert json.loads(srv.stdout.readline()).get("ready")

    def call(*args):
        srv.stdin.write(json.dumps([str(a) for a in args]) + "\n")
        srv.stdin.flush()
        r = json.loads(srv.stdout.readline())
        assert r["ok"], r
        return r["reply"]

    with tempfile.TemporaryDirectory() as tmp:
        for name, kind, w, h, q, rst, il, samp, *rest in CASES:
            seginfo = rest[0] if rest else 0
            if len(sys.argv) > 2 and not name.endswith(sys.argv[2]):
                continue
            jpg, rgb = os.path.join(tmp, "a.jpg"), os.path.join(tmp, "a.rgb")
            extra = ["par:segment_info=1"] if seginfo else []
            if samp == (1, 1):
                call("encode", kind, w, h, q, rst, il, jpg, *extra)
            else:
                call("encode", kind, w, h, q, rst, il, jpg, samp[0], samp[1], *extra)
            call("decode", jpg, rgb)
            np.savez_compressed(os.path.join(out_dir, "refgpu_%s.npz" % name), kind=kind, w=w, h=h, quality=q, rst=rst,
                                interleaved=il, sampling=np.array(samp), segment_info=seginfo, jpeg=np.fromfile(jpg, np.uint8),
                                pixels=np.fromfile(rgb, np.uint8).reshape(h, w, 3))
            print(name, os.path.getsize(jpg), "bytes")
    srv.stdin.close()
    srv.wait(timeout=30)


if __name__ == "__main__":
    main()
