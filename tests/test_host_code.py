"""Host-side code of the product (tables, codestream writer/reader, geometry) against the oracle.
No GPU involved: the C sources are compiled into a shim next to the tests.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import _oracle as o
from _shims import hs


@pytest.mark.parametrize("w,h,q,rst,il", [(1920, 1080, 75, 24, 0), (7680, 4320, 75, 36, 0), (33, 17, 90, 2, 0),
                                          (640, 480, 1, 0, 0), (64, 64, 100, 65535, 1)])
def test_header_bytes_match_oracle(w, h, q, rst, il):
    mine = np.zeros(2048, np.uint8)
    n = hs.shim_header(w, h, q, rst, il, mine)
    want = np.zeros(2048, np.uint8)
    m = o.lib.orc_write_header(want, w, h, q, rst, 3)
    assert np.array_equal(mine[:m], want[:m])
    # followed by the first SOS header: compare against a real oracle stream
    jpeg = o.encode(o.gen_image("zero", 16, 16), q, 1, il)
    info = o.probe(jpeg)
    sos_len = n - m
    assert np.array_equal(mine[m:n], jpeg[info.header_size:info.header_size + sos_len])


@pytest.mark.parametrize("cls", [0, 1])
def test_encoder_lut_matches_oracle_tables(cls):
    ac, dc = np.zeros(256, np.uint32), np.zeros(16, np.uint32)
    hs.shim_enc_lut(cls, ac, dc)
    for kind, lut, n in ((0, dc, 12), (1, ac, 256)):
        code, size = np.zeros(256, np.uint16), np.zeros(256, np.uint8)
        o.lib.orc_huff_encoder_table(cls, kind, code, size)
        assert np.array_equal(lut[:n] & 31, size[:n])
        assert np.array_equal((lut[:n] >> 5)[size[:n] > 0], code[:n][size[:n] > 0])


@pytest.mark.parametrize("cls,kind", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_decoder_lut_decodes_every_code(cls, kind):
    code, size = np.zeros(256, np.uint16), np.zeros(256, np.uint8)
    o.lib.orc_huff_encoder_table(cls, kind, code, size)
    rng = np.random.default_rng(3)
    for sym in range(256):
        if size[sym] == 0:
            continue
        for _ in range(4):
            sz = int(size[sym])
            tail = int(rng.integers(0, 1 << (16 - sz))) if sz < 16 else 0
            peek = (int(code[sym]) << (16 - sz)) | tail
            ln = C.c_int()
            assert hs.shim_dec_lut_symbol(cls, kind, peek, C.byref(ln)) == sym
            assert ln.value == size[sym]


@pytest.mark.parametrize("cls,kind", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_fast_decoder_table_steps_over_every_code(cls, kind):
    """gj_dec_fast (the table of the self-synchronising K3): for every code of the standard tables, followed by random
    bits, one entry gives the bits to consume (code + value), the value size and the zig-zag advance; the standard
    tables never need the canonical search."""
    code, size = np.zeros(256, np.uint16), np.zeros(256, np.uint8)
    o.lib.orc_huff_encoder_table(cls, kind, code, size)
    rng = np.random.default_rng(4)
    levels = set()
    for sym in range(256):
        sz = int(size[sym])
        if sz == 0:
            continue
        vsize, run = sym & 15, sym >> 4
        kadv = 1 if kind == 0 else run + 1 if vsize else 16 if run == 15 else 64
        for _ in range(4):
            tail = int(rng.integers(0, 1 << (16 - sz))) if sz < 16 else 0
            peek = (int(code[sym]) << (16 - sz)) | tail
            total, vs, how = C.c_int(), C.c_int(), C.c_int()
            assert hs.shim_dec_fast_symbol(cls, kind, peek, C.byref(total), C.byref(vs), C.byref(how)) == kadv
            assert (total.value, vs.value) == (sz + vsize, vsize)
            assert how.value == (1 if sz <= 10 else 2)
            levels.add(how.value)
    assert levels == ({1} if (cls, kind) == (0, 0) else {1, 2})   # longest code: DC luminance 9 bits, DC chrominance 11, AC 16


@pytest.mark.parametrize("kind,w,h,rst,il", [("random", 1920, 1080, 24, 0), ("photo", 640, 360, 7, 1),
                                             ("random", 100, 50, 0, 0), ("gradient", 64, 64, 1, 0)])
def test_reader_finds_every_segment(kind, w, h, rst, il):
    jpeg = o.encode(o.gen_image(kind, w, h), 75, rst, il)
    info = np.zeros(8, np.int32)
    cap = 100000
    off, ln = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
    n = hs.shim_parse(jpeg, jpeg.size, info, off, ln, cap)
    p = o.probe(jpeg)
    assert n == p.segment_count
    assert tuple(info[:6]) == (w, h, 3, rst, p.scan_count, il)
    assert info[6] == p.header_size
    # every segment decodes (with the oracle's segment decoder) to the coefficients of the frame
    _, coef = o.decode(jpeg, want_coef=True)
    if not il:
        dw, dh = (w + 7) // 8 * 8, (h + 7) // 8 * 8
        nblk = dw * dh // 64
        seg_mcu = rst if rst else nblk
        per = (nblk + seg_mcu - 1) // seg_mcu
        bits = [[None, None], [None, None]]
        for g in (0, per, 2 * per, n - 1):
            c, s = divmod(g, per)
            cnt = min(seg_mcu, nblk - s * seg_mcu)
            out = np.zeros(cnt * 64, np.int16)
            specs = []
            for k in (0, 1):
                b, v, nv = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint8)(), C.c_int()
                o.lib.orc_huff_spec(0 if c == 0 else 1, k, C.byref(b), C.byref(v), C.byref(nv))
                specs += [b, v]
            o.lib.orc_huff_decode_segment.argtypes = None
            seg = np.ascontiguousarray(jpeg[off[g]:off[g] + ln[g]])
            rc = o.lib.orc_huff_decode_segment(seg.ctypes.data_as(C.c_void_p), C.c_size_t(seg.size), cnt, specs[0],
                                               specs[1], specs[2], specs[3], out.ctypes.data_as(C.c_void_p))
            assert rc == 0
            assert np.array_equal(out, coef[c][s * seg_mcu * 64:(s * seg_mcu + cnt) * 64])


def test_geometry_matches_survey_table():
    out = np.zeros(8, np.int64)
    hs.shim_geometry(7680, 4320, 36, 0, out)
    assert tuple(out[:6]) == (960, 540, 518400, 14400, 43200, 3)
    hs.shim_geometry(1920, 1080, 24, 0, out)
    assert tuple(out[:6]) == (240, 135, 32400, 1350, 4050, 3)
    hs.shim_geometry(1119, 561, 8, 1, out)
    assert tuple(out[:6]) == (140, 71, 9940, 1243, 1243, 1)


@pytest.mark.parametrize("il", [0, 1])
@pytest.mark.parametrize("sampling", [(2, 2), (2, 1), (1, 2), (1, 1)])
@pytest.mark.parametrize("w,h,rst", [(1920, 1080, 12), (1119, 561, 8), (33, 17, 2), (16, 16, 0), (7680, 4320, 6)])
def test_geometry_with_chroma_subsampling(w, h, rst, sampling, il):
    """component planes, scans and restart segments of gj_geometry_init against the reference's rules
    [ref: src/gpujpeg_common.c:671-866], restated independently in tests/_oracle.plane_geometry"""
    out = np.zeros(24, np.int64)
    assert hs.shim_geometry_ss(w, h, rst, il, sampling[0], sampling[1], out) == 0
    planes = o.plane_geometry(w, h, sampling, il)
    off = 0
    for c, (dw, dh) in enumerate(planes):
        assert tuple(out[3 * c:3 * c + 3]) == (dw // 8, dh // 8, off)
        off += dw * dh // 64
    assert out[17] == off * 64
    bpm = sampling[0] * sampling[1] + 2 if il else 1
    if il:
        mcus = [(planes[0][0] // (8 * sampling[0])) * (planes[0][1] // (8 * sampling[1]))]
        assert out[16] == planes[0][0] // (8 * sampling[0])
        # DC predictor distance: 1 inside a component's run of blocks, to the previous MCU otherwise
        n_l = sampling[0] * sampling[1]
        assert out[20] == bpm - (n_l - 1) and all(out[20 + i] == 1 for i in range(1, min(n_l, 4)))
    else:
        mcus = [dw * dh // 64 for dw, dh in planes]
    seg_mcu = rst if rst else max(mcus)
    segs = [(m + seg_mcu - 1) // seg_mcu for m in mcus]
    assert list(out[9:9 + len(segs)]) == segs and out[13] == sum(segs)
    assert out[14] == seg_mcu and out[15] == bpm
    assert out[19] == (1 if sampling == (1, 1) else 0)
    assert out[18] >= seg_mcu * bpm * 416 and out[18] % 128 == 0


@pytest.mark.parametrize("fmt,il,native", [(o.FMT_U8, 0, o.FMT_U8), (o.FMT_444_P012, 1, o.FMT_444_P012),
                                           (o.FMT_444_P012, 0, o.FMT_444_P0P1P2), (o.FMT_422_P1020, 1, o.FMT_422_P1020),
                                           (o.FMT_422_P0P1P2, 0, o.FMT_422_P0P1P2), (o.FMT_420_P0P1P2, 1, o.FMT_420_P0P1P2)])
def test_get_image_info_reports_the_stream_as_the_reference_does(fmt, il, native):
    """gpujpeg_decoder_get_image_info is pure host code (no GPU): size, components, sampling, restart interval,
    interleaving, segment count and the stream's native pixel format [ref: src/gpujpeg_reader.c:1507-1547, 1740-1790]"""
    import ctypes as C
    import gpujpeg_b200.api as api
    w, h, rst = 100, 60, 5
    jpeg = o.encode_ycc(o.gen_raw(fmt, w, h), w, h, fmt, 75, rst, il)
    pi, p, nseg = api.ImageParameters(), api.Parameters(), C.c_int()
    assert api.lib.gpujpeg_decoder_get_image_info(jpeg.ctypes.data, jpeg.size, C.byref(pi), C.byref(p), C.byref(nseg)) == 0
    comps = 1 if fmt == o.FMT_U8 else 3
    assert (pi.width, pi.height, p.comp_count, p.restart_interval) == (w, h, comps, rst)
    assert p.interleaved == (il if comps == 3 else 0) and pi.pixel_format == native
    lh, lv = o.FMT_SAMPLING[fmt]
    assert (p.sampling_factor[0].horizontal, p.sampling_factor[0].vertical) == (lh, lv)
    assert nseg.value == o.probe(jpeg).segment_count


@pytest.mark.parametrize("fmt", range(6))
@pytest.mark.parametrize("w,h", [(64, 48), (1920, 1080), (322, 201), (16, 2)])
def test_raw_layout_matches_the_reference_size_rule_and_the_oracle(fmt, w, h):
    """where the samples of a pixel format live: total size against gpujpeg_image_calculate_size (the reference's own
    rule, src/gpujpeg_common.c:1180-1205) and against the oracle's independent layout; every sample address in range"""
    import ctypes as C
    import gpujpeg_b200.api as api
    out = np.zeros(17, np.int64)
    assert hs.shim_raw_layout(fmt, w, h, 0, out) == 0
    pi = api.image_parameters(w, h, 0, fmt)
    assert out[1] == api.lib.gpujpeg_image_calculate_size(C.byref(pi)) == o.lib.orc_raw_size(fmt, w, h, 0)
    comps = int(out[0])
    assert comps == (1 if fmt == o.FMT_U8 else 3)
    lh, lv = o.FMT_SAMPLING[fmt]
    assert (out[5], out[6]) == (lh, lv)
    for c in range(comps):
        off, pitch, xs = (int(v) for v in out[2 + 5 * c:5 + 5 * c])
        cw = w if c == 0 else (w + lh - 1) // lh
        ch = h if c == 0 else (h + lv - 1) // lv
        assert off + (ch - 1) * pitch + (cw - 1) * xs < out[1]


def test_raw_layout_refuses_what_is_ambiguous():
    out = np.zeros(17, np.int64)
    assert hs.shim_raw_layout(o.FMT_422_P1020, 33, 16, 0, out) == -1      # odd width of packed 4:2:2
    assert hs.shim_raw_layout(o.FMT_420_P0P1P2, 32, 16, 4, out) == -1     # row padding of planar formats
    assert hs.shim_raw_layout(7, 32, 16, 0, out) == -1                     # not a pixel format
    assert hs.shim_raw_layout(6, 32, 16, 0, out) == 0 and out[0] == 3 and out[1] == 4 * 32 * 16   # RGBA: 3 components + alpha
    assert hs.shim_raw_layout(o.FMT_444_P012, 33, 17, 5, out) == 0 and out[3] == 3 * 33 + 5


@pytest.mark.parametrize("internal,fmt,il,samp", [(o.CS_JPEG, o.FMT_U8, 0, (1, 1)), (o.CS_JPEG, o.FMT_444_P012, 1, (2, 2)),
                                                  (o.CS_RGB, o.FMT_444_P012, 0, (1, 1)), (o.CS_RGB, o.FMT_444_P012, 1, (2, 1)),
                                                  (o.CS_601, o.FMT_444_P012, 1, (2, 2)), (o.CS_709, o.FMT_444_P012, 0, (1, 2))])
def test_header_flavours_match_oracle_streams(internal, fmt, il, samp):
    """the host writer's JFIF / Adobe APP14 / SPIFF headers, SOF0 sampling factors, table selectors and the first SOS
    against the first bytes of an oracle stream of the same parameters (the oracle's headers are pinned against the
    reference writer in tests/test_oracle_vs_ref.py)"""
    w, h, q, rst = 70, 50, 80, 3
    comps = 1 if fmt == o.FMT_U8 else 3
    if comps == 1:
        jpeg = o.encode_ycc(o.gen_raw(fmt, w, h), w, h, fmt, q, rst, 0)
    else:
        jpeg = o.encode_any(o.gen_image("photo", w, h), w, h, fmt, o.CS_RGB, q, rst, il, samp, internal=internal)
    mine = np.zeros(4096, np.uint8)
    n = hs.shim_header2(w, h, q, rst, il, comps, samp[0], samp[1], internal, mine)
    assert n > 300 and np.array_equal(mine[:n], jpeg[:n])


class _ImageInfo(__import__("ctypes").Structure):
    """struct gpujpeg_image_info (512 bytes) [ref: libgpujpeg/gpujpeg_decoder.h:243-260]"""
    import ctypes as _C
    import gpujpeg_b200.api as _api
    _fields_ = [("param_image", _api.ImageParameters), ("param", _api.Parameters), ("segment_count", _C.c_int),
                ("header_type", _C.c_int), ("comment", _C.c_char_p), ("pad", _C.c_char * 512)]


def _with_com(jpeg, payload):
    """the stream with a COM segment of `payload` inserted right after SOI"""
    n = len(payload) + 2
    return np.frombuffer(bytes(jpeg[:2]) + b"\xff\xfe" + bytes([n >> 8, n & 255]) + payload + bytes(jpeg[2:]), np.uint8)


def test_com_marker_is_a_c_string_or_nothing_and_ffmpeg_colour_space_comment():
    """COM handling of the reference reader [ref: src/gpujpeg_reader.c:641-672]: the comment is handed out only when
    it is NUL-terminated (callers read it as a C string); FFmpeg's "CS=ITU601" (with or without NUL) names the colour
    space of the components: limited-range BT.601"""
    import ctypes as C
    import gpujpeg_b200.api as api
    fn = api.lib.gpujpeg_decoder_get_image_info2
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(_ImageInfo), C.c_int, C.c_uint]
    base = o.encode(o.gen_image("photo", 64, 48), 75, 4)

    def info(j):
        i = _ImageInfo()
        assert fn(j.ctypes.data, j.size, C.byref(i), 0, 0) == 0
        return i
    own = info(base)
    assert own.comment == b"CREATOR: GPUJPEG, quality = 75" and own.param_image.color_space == api.GPUJPEG_YCBCR_JPEG
    # a later COM without terminator replaces nothing: the first NUL-terminated one stays ... and a stream whose only
    # comment is unterminated has none (the reader must not hand out a pointer into the following file bytes)
    strip = bytes(base).replace(b"\xff\xfe\x00\x21CREATOR: GPUJPEG, quality = 75\x00", b"")
    assert len(strip) < base.size
    j = _with_com(np.frombuffer(strip, np.uint8), b"Lavc60.3.100")
    assert info(j).comment is None
    j = _with_com(np.frombuffer(strip, np.uint8), b"")
    assert info(j).comment is None
    for payload in (b"CS=ITU601", b"CS=ITU601\x00"):
        i = info(_with_com(base, payload))
        assert i.param_image.color_space == api.GPUJPEG_YCBCR_BT601 and i.param.color_space_internal == api.GPUJPEG_YCBCR_BT601
    assert info(_with_com(base, b"CS=ITU6010")).param_image.color_space == api.GPUJPEG_YCBCR_JPEG
