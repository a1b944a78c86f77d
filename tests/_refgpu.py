"""Runs the REFERENCE GPU library (oracle/_ref/libgpujpeg_refgpu.so, compiled in place from /root/reference
by oracle/Makefile `refgpu`) in a process of its own, so its gpujpeg_* symbols never meet the product's.
Test/bench infrastructure only.

    python tests/_refgpu.py encode <kind> <w> <h> <q> <rst> <interleaved> <out.jpg> [<luma_h> <luma_v>]
    python tests/_refgpu.py decode <in.jpg> <out.rgb>
    python tests/_refgpu.py encode_raw <in.raw> <pixfmt> <colorspace> <w> <h> <q> <rst> <interleaved> <out.jpg>
    python tests/_refgpu.py decode_fmt <in.jpg> <colorspace> <pixfmt> <out.raw>
    python tests/_refgpu.py bench  <kind> <w> <h> <q> <rst> <iters>      -> prints JSON with ms per frame
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libgpujpeg_refgpu.so")


class SF(C.Structure):
    _fields_ = [("h", C.c_uint8), ("v", C.c_uint8)]


class Param(C.Structure):
    _fields_ = [("verbose", C.c_int), ("perf_stats", C.c_int), ("quality", C.c_int), ("restart_interval", C.c_int),
                ("interleaved", C.c_int), ("segment_info", C.c_int), ("comp_count", C.c_int), ("sf", SF * 4),
                ("color_space_internal", C.c_int)]


class ImgParam(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("color_space", C.c_int), ("pixel_format", C.c_int),
                ("width_padding", C.c_int)]


class EncIn(C.Structure):
    _fields_ = [("type", C.c_int), ("image", C.c_void_p), ("texture", C.c_void_p)]


class DecOut(C.Structure):
    _fields_ = [("type", C.c_int), ("data", C.c_void_p), ("data_size", C.c_size_t), ("param_image", ImgParam),
                ("texture", C.c_void_p), ("metadata", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("memory_to", "memory_from", "memory_map", "memory_unmap", "preprocessor",
                                          "dct_quantization", "huffman_coder", "stream", "in_gpu")]


def load():
    lib = C.CDLL(SO)
    lib.gpujpeg_encoder_create.restype = C.c_void_p
    lib.gpujpeg_encoder_create.argtypes = [C.c_void_p]
    lib.gpujpeg_decoder_create.restype = C.c_void_p
    lib.gpujpeg_decoder_create.argtypes = [C.c_void_p]
    lib.gpujpeg_encoder_encode.argtypes = [C.c_void_p, C.POINTER(Param), C.POINTER(ImgParam), C.POINTER(EncIn),
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.gpujpeg_decoder_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(DecOut)]
    lib.gpujpeg_encoder_destroy.argtypes = [C.c_void_p]
    lib.gpujpeg_decoder_destroy.argtypes = [C.c_void_p]
    lib.gpujpeg_encoder_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.gpujpeg_decoder_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.gpujpeg_decoder_set_output_format.argtypes = [C.c_void_p, C.c_int, C.c_int]
    assert lib.gpujpeg_init_device(0, 0) == 0
    return lib


def params(lib, w, h, q, rst, il, perf=0):
    p, pi = Param(), ImgParam()
    lib.gpujpeg_set_default_parameters(C.byref(p))
    lib.gpujpeg_image_set_default_parameters(C.byref(pi))
    p.quality, p.restart_interval, p.interleaved, p.perf_stats, p.verbose = q, rst, il, perf, -1
    pi.width, pi.height = w, h
    return p, pi


def gen(kind, w, h):
    sys.path.insert(0, HERE)
    import _oracle as o
    return o.gen_image(kind, w, h)


def encode(lib, enc, img, p, pi):
    inp = EncIn(0, img.ctypes.data, None)
    out, size = C.c_void_p(), C.c_size_t()
    rc = lib.gpujpeg_encoder_encode(enc, C.byref(p), C.byref(pi), C.byref(inp), C.byref(out), C.byref(size))
    assert rc == 0, rc
    return np.ctypeslib.as_array((C.c_uint8 * size.value).from_address(out.value))


def main():
    mode = sys.argv[1]
    lib = load()
    if mode == "encode":
        kind, w, h, q, rst, il, path = sys.argv[2], *map(int, sys.argv[3:8]), sys.argv[8]
        img = gen(kind, w, h)
        enc = lib.gpujpeg_encoder_create(None)
        p, pi = params(lib, w, h, q, rst, il)
        if len(sys.argv) > 10:   # chroma subsampling: GPUJPEG_SUBSAMPLING_xxx packing, first component on top
            lib.gpujpeg_parameters_chroma_subsampling.argtypes = [C.POINTER(Param), C.c_uint32]
            lib.gpujpeg_parameters_chroma_subsampling(C.byref(p), int(sys.argv[9]) << 28 | int(sys.argv[10]) << 24 | 0x111100)
        encode(lib, enc, img, p, pi).tofile(path)
        lib.gpujpeg_encoder_destroy(enc)
    elif mode == "encode_raw":
        src, fmt, cs, w, h, q, rst, il, path = sys.argv[2], *map(int, sys.argv[3:10]), sys.argv[10]
        raw = np.fromfile(src, np.uint8)
        enc = lib.gpujpeg_encoder_create(None)
        p, pi = params(lib, w, h, q, rst, il)
        pi.pixel_format, pi.color_space = fmt, cs   # comp_count stays 0: sampling follows the pixel format
        if len(sys.argv) > 11:
            p.color_space_internal = int(sys.argv[11])   # e.g. 1 = GPUJPEG_RGB: RGB-internal JPEG (Adobe APP14)
        encode(lib, enc, raw, p, pi).tofile(path)
        lib.gpujpeg_encoder_destroy(enc)
    elif mode == "decode_fmt":
        data = np.fromfile(sys.argv[2], np.uint8)
        dec = lib.gpujpeg_decoder_create(None)
        lib.gpujpeg_decoder_set_output_format(dec, int(sys.argv[3]), int(sys.argv[4]))
        out = DecOut()
        out.type = 0
        assert lib.gpujpeg_decoder_decode(dec, data.ctypes.data, data.size, C.byref(out)) == 0
        np.ctypeslib.as_array((C.c_uint8 * out.data_size).from_address(out.data)).tofile(sys.argv[5])
        print(json.dumps({"pixel_format": out.param_image.pixel_format, "color_space": out.param_image.color_space,
                          "size": out.data_size}))
        lib.gpujpeg_decoder_destroy(dec)
    elif mode == "decode":
        data = np.fromfile(sys.argv[2], np.uint8)
        dec = lib.gpujpeg_decoder_create(None)
        lib.gpujpeg_decoder_set_output_format(dec, 1, 1)  # GPUJPEG_RGB, GPUJPEG_444_U8_P012
        out = DecOut()
        out.type = 0
        assert lib.gpujpeg_decoder_decode(dec, data.ctypes.data, data.size, C.byref(out)) == 0
        np.ctypeslib.as_array((C.c_uint8 * out.data_size).from_address(out.data)).tofile(sys.argv[3])
        lib.gpujpeg_decoder_destroy(dec)
    elif mode == "bench":
        kind, w, h, q, rst, iters = sys.argv[2], *map(int, sys.argv[3:8])
        img = gen(kind, w, h)
        enc = lib.gpujpeg_encoder_create(None)
        dec = lib.gpujpeg_decoder_create(None)
        lib.gpujpeg_decoder_set_output_format(dec, 1, 1)
        p, pi = params(lib, w, h, q, rst, 0, perf=1)
        jpeg = encode(lib, enc, img, p, pi).copy()
        te, tg = [], []
        for _ in range(iters):
            t = time.perf_counter()
            encode(lib, enc, img, p, pi)
            te.append((time.perf_counter() - t) * 1e3)
            s = Stats()
            if lib.gpujpeg_encoder_get_stats(enc, C.byref(s)) == 0:
                tg.append(s.in_gpu)
        td, tdg = [], []
        out = DecOut()
        for _ in range(iters):
            out.type = 0
            t = time.perf_counter()
            assert lib.gpujpeg_decoder_decode(dec, jpeg.ctypes.data, jpeg.size, C.byref(out)) == 0
            td.append((time.perf_counter() - t) * 1e3)
            s = Stats()
            if lib.gpujpeg_decoder_get_stats(dec, C.byref(s)) == 0:
                tdg.append(s.in_gpu)
        med = lambda x: float(np.median(x)) if len(x) else None
        print(json.dumps({"impl": "reference-gpu", "w": w, "h": h, "kind": kind, "jpeg_bytes": int(jpeg.size),
                          "encode_ms_e2e": med(te), "encode_ms_gpu": med(tg), "decode_ms_e2e": med(td),
                          "decode_ms_gpu": med(tdg), "iters": iters}))


if __name__ == "__main__":
    main()
