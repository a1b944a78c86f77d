"""The kernels' per-thread arithmetic (gpujpeg_b200/csrc/gj_device.cuh) compiled for the host and
checked against the oracle -- catches arithmetic bugs in the build container, where no GPU exists.
The GPU parity tests (test_gpu_*.py) then only have to prove the memory/indexing code.  CPU only."""
import numpy as np
import pytest

import _oracle as o
from _shims import hs, km


def test_zigzag_tables():
    assert km.km_zigzag_tables_consistent() == 1


def test_float_colour_transform_is_exact_for_all_inputs():
    # 2^24 RGB triples: the float evaluation in K1 must equal the reference's integer arithmetic
    assert km.km_check_rgb_to_ycbcr_exhaustive(0, 256) == 0


def test_inverse_colour_transform_is_exact_for_all_inputs():
    assert km.km_check_ycbcr_to_rgb_exhaustive(0, 256) == 0


@pytest.mark.parametrize("kind,q", [("random", 75), ("photo", 75), ("random", 100), ("gradient", 10), ("photo", 95)])
def test_fdct_block_matches_oracle(kind, q):
    w, h = 512, 256
    img = o.gen_image(kind, w, h)
    planes = np.zeros((3, h * w), np.uint8)
    o.lib.orc_preprocess_rgb444(img.reshape(-1), w, h, 0, planes.reshape(-1), w, h)
    _, fwd, _ = o.quant_tables(q)
    for c in range(3):
        cls = 0 if c == 0 else 1
        want = np.zeros(w * h, np.int16)
        o.lib.orc_fdct_quant_plane(planes[c], w, h, fwd[cls], want)
        fwd_zz, raw = np.zeros(64, np.float32), np.zeros(64, np.uint8)
        hs.shim_forward_table_zz(cls, q, fwd_zz, raw)
        got = np.zeros(w * h, np.int16)
        km.km_fdct_quant_plane(planes[c], w, h, fwd_zz, got)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("flavour", [o.IDCT_INT, o.IDCT_FLOAT_GPUREF])
@pytest.mark.parametrize("kind,q", [("random", 75), ("photo", 75), ("random", 100), ("gradient", 25)])
def test_idct_block_matches_oracle(kind, q, flavour):
    w, h = 256, 256
    img = o.gen_image(kind, w, h)
    _, coef = o.encode(img, q, 8, want_coef=True)
    raw, _, inv = o.quant_tables(q)
    for c in range(3):
        cls = 0 if c == 0 else 1
        want = np.zeros(w * h, np.uint8)
        o.lib.orc_idct_plane(np.ascontiguousarray(coef[c]), w, h, inv[cls], flavour, want)
        got = np.zeros(w * h, np.uint8)
        km.km_idct_plane(np.ascontiguousarray(coef[c]), w, h, raw[cls].astype(np.uint16), flavour, got)
        assert np.array_equal(got, want)


def test_huffman_value_helpers():
    for v in list(range(-2047, 2048)):
        n = km.km_category(v)
        assert n == (abs(v)).bit_length()
        if v:
            bits = km.km_value_bits(v, n)
            assert km.km_extend(bits, n) == v
