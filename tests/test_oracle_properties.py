"""Self-consistency of the oracle: independent decoder (PIL/libjpeg) agreement, PSNR thresholds the
reference's regression suite uses (test/regression/run_tests.sh:116-151), colour-transform identities.
CPU only."""
import io

import numpy as np
import pytest

import _oracle as o

PIL = pytest.importorskip("PIL.Image")


def psnr(a, b):
    mse = ((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean()
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def test_reference_psnr_threshold_random_rgb():
    # reference: 1119x561 LCG-random RGB q75 must round-trip at >= 22 dB (measured 22.26)
    img = o.gen_image("random", 1119, 561)
    jpeg = o.encode(img, 75, 8)
    assert psnr(o.decode(jpeg), img) >= 22.0
    pil = np.array(PIL.open(io.BytesIO(jpeg.tobytes())).convert("RGB"))
    assert psnr(pil, img) >= 22.0
    assert np.abs(pil.astype(int) - o.decode(jpeg).astype(int)).max() <= 4


def test_zero_image_round_trips_exactly():
    # reference test_commit_b620be2: all-zero image must round-trip (>= 50 dB)
    img = o.gen_image("zero", 1920, 1080)
    dec = o.decode(o.encode(img, 75, 24))
    assert psnr(dec, img) >= 50.0


@pytest.mark.parametrize("kind,rst,il", [("photo", 24, 0), ("random", 5, 1), ("gradient", 0, 0)])
def test_pil_decodes_stream(kind, rst, il):
    img = o.gen_image(kind, 320, 200)
    jpeg = o.encode(img, 75, rst, il)
    pil = np.array(PIL.open(io.BytesIO(jpeg.tobytes())).convert("RGB"))
    assert pil.shape == img.shape
    assert np.abs(pil.astype(int) - o.decode(jpeg).astype(int)).max() <= 4


def test_colour_scaling_identities():
    # (c*256)/255 == c + (c==255) for c in 0..255 ; ((v)*256)/255 == v for |v| <= 128 (C truncation)
    c = np.arange(256)
    assert np.array_equal(c * 256 // 255, c + (c == 255))
    v = np.arange(-128, 128)
    assert np.array_equal(np.trunc(v * 256 / 255).astype(int), v)


def test_threads_do_not_change_bytes():
    img = o.gen_image("photo", 640, 360)
    assert np.array_equal(o.encode(img, 75, 24, threads=1), o.encode(img, 75, 24, threads=4))
    j = o.encode(img, 75, 24)
    assert np.array_equal(o.decode(j, threads=1), o.decode(j, threads=4))


def test_probe_counts_segments():
    img = o.gen_image("random", 1920, 1080)
    info = o.probe(o.encode(img, 75, 24))
    assert (info.width, info.height, info.comp_count, info.restart_interval) == (1920, 1080, 3, 24)
    assert info.scan_count == 3 and info.segment_count == 4050
