"""Full-size frames of BASELINE.json (8K 4:4:4, 8K 4:2:0, 16K): product vs oracle through the C ABI.

Kept in their own module, named so that it is collected LAST: `pytest -x` then reports every other GPU row first.
Memory is bounded: the coder pair of this module is created per test and destroyed afterwards (a 16K encoder
holds its worst-case scan buffers only while it lives), the oracle runs on 4 threads, and the big arrays are
dropped before the next test starts."""
import gc

import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu
ORACLE_THREADS = 4


@pytest.fixture()
def coders():
    import gpujpeg_b200 as gj
    e, d = gj.Encoder(), gj.Decoder()
    yield e, d
    e.close()
    d.close()
    gc.collect()

def test_full_size_8k_round_trip(coders):
    """BASELINE configs 3+4 at full size: bytes vs the (multi-threaded) oracle, then decode parity."""
    enc, dec = coders
    w, h = 7680, 4320
    img = o.gen_image("photo", w, h)
    want = o.encode(img, 75, 36, threads=ORACLE_THREADS)
    got = enc.encode(img, 75, 36)
    assert got.size == want.size and np.array_equal(got, want)
    out = dec.decode(got)
    assert np.array_equal(out, o.decode(want, threads=ORACLE_THREADS))
    info = o.probe(got)
    assert info.segment_count == 43200

def test_subsampled_8k_round_trip(coders):
    """full-size 4:2:0 interleaved (what video pipelines feed): bytes and pixels against the threaded oracle,
    then back to 4:4:4 on the same coder instances (re-initialisation across sampling modes)"""
    enc, dec = coders
    w, h = 7680, 4320
    img = o.gen_image("photo", w, h)
    want = o.encode(img, 75, 6, 1, threads=ORACLE_THREADS, sampling=(2, 2))
    got = enc.encode(img, 75, 6, 1, subsampling="4:2:0")
    assert got.size == want.size and np.array_equal(got, want)
    assert np.array_equal(dec.decode(got), o.decode(want, threads=ORACLE_THREADS))
    small = o.gen_image("random", 64, 64)
    j = enc.encode(small, 75, 4)
    assert np.array_equal(j, o.encode(small, 75, 4))
    assert np.array_equal(dec.decode(j), o.decode(j))

def test_full_size_16k_round_trip(coders):
    """BASELINE config 5 frame size (15360x8640, RESTART_AUTO = 36): 172 800 segments, 6.2 M blocks"""
    enc, dec = coders
    w, h = 15360, 8640
    img = o.gen_image("photo", w, h)
    want = o.encode(img, 75, 36, threads=ORACLE_THREADS)
    got = enc.encode(img, 75, 36)
    assert got.size == want.size and np.array_equal(got, want)
    assert np.array_equal(dec.decode(got), o.decode(want, threads=ORACLE_THREADS))
    assert o.probe(got).segment_count == 172800
