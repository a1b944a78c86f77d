"""Foreign-stream robustness of the decoder: streams that carry the same image as one
of the codec family's own files but are laid out the way other encoders do it -- other table ids, no DRI marker,
extra marker segments, fill bytes -- must decode to the same pixels; streams outside baseline JPEG or with a broken
restart structure must be refused with an error, never decoded to garbage or crash.  GPU, through the C ABI."""
import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gj():
    import gpujpeg_b200
    return gpujpeg_b200


@pytest.fixture(scope="module")
def dec(gj):
    d = gj.Decoder()
    yield d
    d.close()


def segments(j):
    """[(marker, start, end)] of the marker segments in front of the first SOS, plus the SOS position"""
    j = bytes(j)
    i, out = 2, []
    while j[i + 1] != 0xDA:
        ln = (j[i + 2] << 8) | j[i + 3]
        out.append((j[i + 1], i, i + 2 + ln))
        i += 2 + ln
    return out, i


def swap_table_ids(jpeg):
    """quantisation tables 0<->1, DC Huffman tables 0<->1 (AC stay): luminance then uses Tq 1, Td 1, Ta 0"""
    j = bytearray(jpeg)
    segs, _ = segments(j)
    for m, a, b in segs:
        if m == 0xDB:
            j[a + 4] ^= 1
        elif m == 0xC4 and (j[a + 4] >> 4) == 0:
            j[a + 4] ^= 1
        elif m == 0xC0:
            for c in range(j[a + 9]):
                j[a + 10 + 3 * c + 2] ^= 1
    i = 0
    while True:   # every SOS: flip the DC selector of each component
        i = bytes(j).find(b"\xff\xda", i)
        if i < 0:
            break
        for c in range(j[i + 4]):
            j[i + 5 + 2 * c + 1] ^= 0x10
        i += 4
    return np.frombuffer(bytes(j), np.uint8)


CASES = [("photo", 320, 200, 80, 6, 0, (1, 1)), ("random", 161, 97, 90, 4, 1, (2, 2)), ("photo", 128, 64, 75, 0, 1, (2, 1))]


@pytest.mark.parametrize("kind,w,h,q,rst,il,samp", CASES)
def test_other_table_ids(dec, kind, w, h, q, rst, il, samp):
    jpeg = o.encode(o.gen_image(kind, w, h), q, rst, il, sampling=samp)
    want = o.decode(jpeg)
    other = swap_table_ids(jpeg)
    assert not np.array_equal(other, jpeg)
    assert np.array_equal(o.decode(other), want), "oracle must read the re-labelled stream as the same image"
    assert np.array_equal(dec.decode(other), want)


@pytest.mark.parametrize("il,samp", [(0, (1, 1)), (1, (2, 2))])
def test_stream_without_dri_and_with_extra_segments(dec, il, samp):
    """no DRI marker at all (restart interval 0 by default), an Exif-like APP1, an APP13, two comments"""
    jpeg = bytes(o.encode(o.gen_image("photo", 200, 120), 85, 0, il, sampling=samp))
    want = o.decode(np.frombuffer(jpeg, np.uint8))
    segs, sos = segments(jpeg)
    dri = [(a, b) for m, a, b in segs if m == 0xDD][0]
    extra = b"\xff\xe1\x00\x10Exif\x00\x00MM\x00\x2a\x00\x00\x00\x08" + b"\xff\xed\x00\x06abcd" + b"\xff\xfe\x00\x05hi\x00"
    other = jpeg[:dri[0]] + extra + jpeg[dri[1]:sos] + b"\xff\xfe\x00\x04ok" + jpeg[sos:]
    other = np.frombuffer(other, np.uint8)
    assert np.array_equal(o.decode(other), want)
    assert np.array_equal(dec.decode(other), want)


def test_fill_bytes_before_restart_markers(dec):
    jpeg = bytes(o.encode(o.gen_image("random", 96, 64), 75, 3, 0))
    want = o.decode(np.frombuffer(jpeg, np.uint8))
    _, sos = segments(jpeg)
    body = bytearray()
    i = sos
    while i < len(jpeg):
        if jpeg[i] == 0xFF and 0xD0 <= jpeg[i + 1] <= 0xD7 and (i // 7) % 2 == 0:
            body += b"\xff\xff"          # one fill byte in front of about half of the RSTn markers
            body.append(jpeg[i + 1])
            i += 2
        else:
            body.append(jpeg[i])
            i += 1
    other = np.frombuffer(jpeg[:sos] + bytes(body), np.uint8)
    assert other.size > len(jpeg)
    assert np.array_equal(o.decode(other), want)
    assert np.array_equal(dec.decode(other), want)


def test_tables_merged_into_one_segment_each(dec):
    """all DQT tables in one DQT segment and all DHT tables in one DHT segment (what libjpeg writes with -optimize off)"""
    jpeg = bytes(o.encode(o.gen_image("photo", 160, 96), 75, 5, 1))
    want = o.decode(np.frombuffer(jpeg, np.uint8))
    segs, sos = segments(jpeg)
    dqt = b"".join(jpeg[a + 4:b] for m, a, b in segs if m == 0xDB)
    dht = b"".join(jpeg[a + 4:b] for m, a, b in segs if m == 0xC4)
    rest = b"".join(jpeg[a:b] for m, a, b in segs if m not in (0xDB, 0xC4))
    mk = lambda code, payload: bytes([0xFF, code, (len(payload) + 2) >> 8, (len(payload) + 2) & 255]) + payload
    other = np.frombuffer(jpeg[:2] + mk(0xDB, dqt) + rest + mk(0xC4, dht) + jpeg[sos:], np.uint8)
    assert np.array_equal(o.decode(other), want)
    assert np.array_equal(dec.decode(other), want)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("kind,w,h,q,rst,il,samp", [("photo", 320, 200, 80, 6, 0, (1, 1)), ("random", 161, 97, 95, 4, 1, (2, 2))])
def test_custom_huffman_tables(dec, seed, kind, w, h, q, rst, il, samp):
    """four random Huffman tables (code lengths up to 16, not Annex K) as an optimising encoder would write them: the
    decoder's lookup tables are built from the DHT segments of the stream, whatever they hold"""
    img = o.gen_image(kind, w, h)
    want = o.decode(o.encode(img, q, rst, il, sampling=samp))
    with o.huffman_override(np.random.default_rng(seed)):
        other = o.encode(img, q, rst, il, sampling=samp)
    assert not np.array_equal(other, o.encode(img, q, rst, il, sampling=samp)), "override must change the stream"
    assert np.array_equal(o.decode(other), want), "same coefficients, other codes: same pixels"
    assert np.array_equal(dec.decode(other), want)


def expect_error(gj, data):
    d = gj.Decoder()
    try:
        with pytest.raises(gj.GpuJpegError):
            d.decode(np.frombuffer(bytes(data), np.uint8))
    finally:
        d.close()


def test_refused_streams(gj):
    jpeg = bytearray(o.encode(o.gen_image("photo", 128, 96), 75, 4, 0))
    segs, sos = segments(jpeg)
    # progressive frame header
    sof = [a for m, a, b in segs if m == 0xC0][0]
    prog = bytearray(jpeg)
    prog[sof + 1] = 0xC2
    expect_error(gj, prog)
    # 12-bit samples
    deep = bytearray(jpeg)
    deep[sof + 4] = 12
    expect_error(gj, deep)
    # 16-bit quantisation table
    dqt = [a for m, a, b in segs if m == 0xDB][0]
    wide = bytearray(jpeg)
    wide[dqt + 4] |= 0x10
    expect_error(gj, wide)
    # truncated in the middle of the scan data, and in the middle of the headers
    expect_error(gj, jpeg[:sos + 40])
    expect_error(gj, jpeg[:sof + 6])
    # not a JPEG at all
    expect_error(gj, b"\x89PNG\r\n\x1a\n" + bytes(64))


def test_decoder_survives_garbage_entropy_data(gj, dec):
    """random bytes in place of the entropy-coded data must not crash or hang the GPU; the call may succeed or fail"""
    jpeg = bytearray(o.encode(o.gen_image("photo", 256, 128), 75, 8, 0))
    _, sos = segments(jpeg)
    rng = np.random.default_rng(9)
    end = len(jpeg) - 2
    for i in range(sos + 12, end):
        if not (jpeg[i] == 0xFF or jpeg[i - 1] == 0xFF):     # keep the marker structure, scramble the rest
            jpeg[i] = int(rng.integers(0, 255))
    try:
        out = dec.decode(np.frombuffer(bytes(jpeg), np.uint8))
        assert out.shape == (128, 256, 3)
    except gj.GpuJpegError:
        pass
    good = o.encode(o.gen_image("photo", 64, 64), 75, 4)
    assert np.array_equal(dec.decode(good), o.decode(good)), "the coder must stay usable afterwards"


def test_scans_in_another_component_order_are_refused(gj):
    """one scan per component, but the first scan codes Cb: the per-scan block counts of the geometry would no longer
    belong to the plane the scan is written to (with 4:2:0 that was an out-of-bounds write) -- refused, not decoded"""
    jpeg = o.encode(o.gen_image("photo", 64, 48), 75, 0, 0, sampling=(2, 2))
    j = bytearray(jpeg)
    sos = [i for i in range(len(j) - 1) if j[i] == 0xFF and j[i + 1] == 0xDA]
    assert len(sos) == 3
    # swap the component selectors (and table selectors) of the first two scan headers
    a, b = sos[0], sos[1]
    j[a + 5], j[b + 5] = j[b + 5], j[a + 5]
    j[a + 6], j[b + 6] = j[b + 6], j[a + 6]
    d = gj.Decoder()
    try:
        with pytest.raises(gj.GpuJpegError):
            d.decode(np.frombuffer(bytes(j), np.uint8))
        assert np.array_equal(d.decode(jpeg), o.decode(jpeg))   # the decoder instance is still usable
    finally:
        d.close()


def test_ffmpeg_cs_itu601_comment_selects_limited_range(gj):
    """a JFIF stream carrying FFmpeg's COM "CS=ITU601" holds limited-range BT.601 samples
    [ref: src/gpujpeg_reader.c:655-661]: the RGB the decoder returns is the BT.601 conversion, and the same stream
    read with ff_cs_itu601_is_709 is converted with the BT.709 matrix"""
    w, h = 64, 48
    raw = o.gen_raw(o.FMT_444_P012, w, h)
    jpeg = o.encode_any(raw, w, h, o.FMT_444_P012, o.CS_601, 90, 4, 1, (1, 1), internal=o.CS_601)
    assert b"CS=ITU601" in bytes(jpeg)
    # strip the SPIFF header so that only the comment names the colour space: re-wrap the scan under a JFIF header
    own = o.encode_any(raw, w, h, o.FMT_444_P012, o.CS_JPEG, 90, 4, 1, (1, 1))
    body = bytes(own)
    com = b"\xff\xfe\x00\x0cCS=ITU601\x00"
    tagged = np.frombuffer(body[:2] + body[2:20] + com + body[20:], np.uint8)   # after the 18-byte JFIF APP0
    d = gj.Decoder()
    try:
        d.set_output_format(gj.api.GPUJPEG_YCBCR_BT601, o.FMT_444_P012)
        out, pi = d.decode_samples(tagged)          # asked for BT.601: the samples come back untransformed
        d.set_output_format(gj.api.GPUJPEG_YCBCR_JPEG, o.FMT_444_P012)
        plain, _ = d.decode_samples(own)            # the same scan read as full-range YCbCr, asked for full-range
        assert np.array_equal(out, plain)
    finally:
        d.close()


@pytest.mark.parametrize("huffman", ["auto", "thread_per_segment"])
@pytest.mark.parametrize("kind,w,h,rst,il,samp", [("photo", 256, 192, 4, 0, (1, 1)), ("random", 200, 120, 3, 1, (1, 1)), ("photo", 320, 200, 2, 1, (2, 2))])
def test_broken_restart_sequences_are_resynchronised_like_the_reference(gj, huffman, kind, w, h, rst, il, samp):
    """[ref: src/gpujpeg_reader.c:1038-1155] a restart marker with the wrong number ends the current segment, the data up to
    the next marker with the EXPECTED number is skipped, later segments move up and the last ones are missing (zero
    blocks); a marker that is missing altogether is the same case.  The decoder must give exactly the picture the
    reference's reader + decoder give -- restated in the oracle's stream splitter."""
    jpeg = bytearray(o.encode(o.gen_image(kind, w, h), 80, rst, il, sampling=samp))
    sos = bytes(jpeg).find(b"\xff\xda")
    marks = [i for i in range(sos, len(jpeg) - 1) if jpeg[i] == 0xFF and 0xD0 <= jpeg[i + 1] <= 0xD7]
    assert len(marks) > 20
    d = gj.Decoder()
    d.set_option("dec_opt_huffman", huffman)
    try:
        # (1) one marker carries the wrong number
        bad = bytearray(jpeg)
        bad[marks[5] + 1] = 0xD0 + ((bad[marks[5] + 1] - 0xD0 + 3) & 7)
        bad = np.frombuffer(bytes(bad), np.uint8)
        want = o.decode(bad)
        assert not np.array_equal(want, o.decode(np.frombuffer(bytes(jpeg), np.uint8)))
        assert np.array_equal(d.decode(bad), want), "wrong-number marker: picture differs from the reference's resynchronisation"
        # (2) one marker is missing: two segments run into each other, the count no longer fits the geometry
        cut = np.frombuffer(bytes(jpeg[:marks[7]] + jpeg[marks[7] + 2:]), np.uint8)
        want = o.decode(cut)
        got = d.decode(cut)
        # the merged segment decodes its own blocks and then runs out of blocks: everything else is well defined
        assert got.shape == want.shape
        seg_rows = np.any(got != want, axis=(1, 2)).sum()
        assert seg_rows <= 16 * max(1, samp[1]), "missing marker: more than the damaged segment differs (%d rows)" % seg_rows
        # (3) the undamaged stream still decodes exactly on the same decoder
        ok = np.frombuffer(bytes(jpeg), np.uint8)
        assert np.array_equal(d.decode(ok), o.decode(ok))
    finally:
        d.close()


def test_image_range_info_prints_the_sample_ranges(gj, tmp_path, capfd):
    """gpujpeg_image_range_info [ref: src/gpujpeg_common.c:1383-1442]: smallest and largest sample per component of a raw
    file, packed 4:4:4 and UYVY (whose even pixels' chrominance byte is filed under component 3, odd pixels' under 2)"""
    import ctypes as C
    lib = gj.api.lib
    lib.gpujpeg_image_range_info.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.gpujpeg_image_range_info.restype = None
    rgb = np.zeros((4, 6, 3), np.uint8)
    rgb[..., 0], rgb[..., 1], rgb[..., 2] = 10, 20, 30
    rgb[1, 2], rgb[3, 5] = (7, 99, 30), (200, 20, 1)
    path = tmp_path / "a.rgb"
    rgb.tofile(path)
    lib.gpujpeg_image_range_info(str(path).encode(), 6, 4, 1)      # GPUJPEG_444_U8_P012
    out = capfd.readouterr().out
    assert "Component 1: 7 - 200" in out and "Component 2: 20 - 99" in out and "Component 3: 1 - 30" in out
    uyvy = np.array([[50, 16, 60, 235, 40, 100, 90, 101]], np.uint8)   # U0 Y0 V0 Y1 | U1 Y2 V1 Y3
    path = tmp_path / "b.uyvy"
    uyvy.tofile(path)
    lib.gpujpeg_image_range_info(str(path).encode(), 4, 1, 3)      # GPUJPEG_422_U8_P1020
    out = capfd.readouterr().out
    assert "Component 1: 16 - 235" in out and "Component 2: 60 - 90" in out and "Component 3: 40 - 50" in out
