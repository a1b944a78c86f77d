"""Pins the CPU oracle against golden vectors produced by the REFERENCE's own code.  CPU only.
  tests/golden/<name>.npz          the reference's CPU sources (Huffman coders, integer IDCT, writer), generated here by
                                   tests/golden/make_golden.py from oracle/_ref
  tests/golden/refgpu_<name>.npz   the reference's GPU library on a B200 (tests/golden/make_golden_refgpu.py): JPEG bytes of
                                   its encoder and pixels of its decoder -- what pins the oracle's restated colour
                                   transforms, float FDCT and float IDCT without a GPU"""
import glob
import os

import numpy as np
import pytest

import _oracle as o

ALL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLDEN = [p for p in ALL if not os.path.basename(p).startswith("refgpu_")]
REFGPU = [p for p in ALL if os.path.basename(p).startswith("refgpu_")]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    w, h, q, rst, il = int(g["w"]), int(g["h"]), int(g["quality"]), int(g["rst"]), int(g["interleaved"])
    img = o.gen_image(str(g["kind"]), w, h)
    if "sampling" in g:   # chroma-subsampled fixtures: bytes and decoded coefficients from the reference CPU code
        samp = tuple(int(v) for v in g["sampling"])
        jpeg, coef = o.encode(img, q, rst, il, want_coef=True, sampling=samp)
        assert np.array_equal(coef, g["coef"])
        assert np.array_equal(jpeg, g["jpeg"]), "oracle JPEG bytes differ from reference writer+CPU Huffman encoder"
        rgb, coef_dec = o.decode(g["jpeg"], o.IDCT_INT, want_coef=True)
        assert np.array_equal(coef_dec, g["coef_dec"]), "oracle Huffman decode differs from reference CPU decoder"
        assert rgb.shape == (h, w, 3)
        return
    jpeg, coef = o.encode(img, q, rst, il, want_coef=True)
    # coefficients are the oracle's own restatement (regression pin), bytes are the reference's
    assert np.array_equal(coef, g["coef"])
    assert np.array_equal(jpeg, g["jpeg"]), "oracle JPEG bytes differ from reference writer+CPU Huffman encoder"
    rgb, coef_dec = o.decode(g["jpeg"], o.IDCT_INT, want_coef=True)
    assert np.array_equal(coef_dec, g["coef_dec"]), "oracle Huffman decode differs from reference CPU decoder"
    # integer IDCT planes: re-run the oracle plane IDCT on the reference-decoded coefficients
    _, _, inv = o.quant_tables(q)
    dw, dh = (w + 7) // 8 * 8, (h + 7) // 8 * 8
    for c in range(3):
        plane = np.zeros(dw * dh, np.uint8)
        o.lib.orc_idct_plane(np.ascontiguousarray(g["coef_dec"][c]), dw, dh, inv[0 if c == 0 else 1], o.IDCT_INT, plane)
        assert np.array_equal(plane.reshape(dh, dw), g["planes"][c]), "oracle integer IDCT differs from gpujpeg_idct_cpu"
    assert rgb.shape == (h, w, 3)


def test_golden_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", REFGPU, ids=[os.path.basename(p)[7:-4] for p in REFGPU])
def test_oracle_matches_reference_gpu_library_outputs(path):
    """the encode front half (colour transform + float FDCT + quantisation) and the float IDCT have no CPU code in the
    reference; these bytes and pixels were produced by its CUDA kernels on a B200"""
    import hashlib
    g = np.load(path)
    w, h, q, rst, il = int(g["w"]), int(g["h"]), int(g["quality"]), int(g["rst"]), int(g["interleaved"])
    samp = tuple(int(v) for v in g["sampling"])
    img = o.gen_image(str(g["kind"]), w, h)
    if "segment_info" in g and int(g["segment_info"]):
        with o.segment_info():
            jpeg = o.encode(img, q, rst, il, sampling=samp)
    else:
        jpeg = o.encode(img, q, rst, il, sampling=samp)
    assert jpeg.size == g["jpeg"].size and np.array_equal(jpeg, g["jpeg"]), "oracle JPEG bytes differ from the reference GPU encoder"
    rgb = o.decode(g["jpeg"], o.IDCT_FLOAT_GPUREF)
    if "pixels" in g:
        assert np.array_equal(rgb, g["pixels"]), "oracle float-IDCT decode differs from the reference GPU decoder"
    else:
        assert hashlib.sha256(np.ascontiguousarray(rgb).tobytes()).hexdigest() == str(g["pixels_sha256"])


def test_reference_gpu_fixtures_present():
    assert len(REFGPU) >= 10
