"""Runs the REFERENCE GPU library (oracle/_ref/libgpujpeg_refgpu.so, compiled in place from /root/reference
by oracle/Makefile `refgpu`) in a process of its own, so its gpujpeg_* symbols never meet the product's.
Test/bench infrastructure only.

    python tests/_refgpu.py encode <kind> <w> <h> <q> <rst> <interleaved> <out.jpg> [<luma_h> <luma_v>]
    python tests/_refgpu.py decode <in.jpg> <out.rgb>
    python tests/_refgpu.py encode_raw <in.raw> <pixfmt> <colorspace> <w> <h> <q> <rst> <interleaved> <out.jpg>
    python tests/_refgpu.py decode_fmt <in.jpg> <colorspace> <pixfmt> <out.raw>
    python tests/_refgpu.py bench  <kind> <w> <h> <q> <rst> <iters>      -> prints JSON with ms per frame
    python tests/_refgpu.py serve        one JSON list of the arguments above per stdin line, one JSON reply per line
                                         (one process, one CUDA context for a whole test session)
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


def pinned_copy(lib, arr):
    """a copy of `arr` in page-locked host memory obtained exactly as the reference's own harness obtains it:
    written to a scratch file and read back by the reference library's gpujpeg_image_load_from_file
    (cudaMallocHost, src/gpujpeg_common.c:1217-1254; gpujpegtool does the same, src/main.c:786-797).
    Returns (numpy view, pinned?, owner address for gpujpeg_image_destroy)."""
    import tempfile
    lib.gpujpeg_image_load_from_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    with tempfile.NamedTemporaryFile(suffix=".rgb") as f:
        np.ascontiguousarray(arr).reshape(-1).tofile(f.name)
        ptr, size = C.c_void_p(), C.c_size_t(0)
        if lib.gpujpeg_image_load_from_file(f.name.encode(), C.byref(ptr), C.byref(size)) != 0 or size.value != arr.size:
            return np.ascontiguousarray(arr).reshape(-1).copy(), False, None
    return np.ctypeslib.as_array((C.c_uint8 * size.value).from_address(ptr.value)), True, ptr


def main():
    mode = sys.argv[1]
    lib = load()
    if mode == "serve":
        sys.stdout.write(json.dumps({"ready": True}) + "\n")
        sys.stdout.flush()
        for line in sys.stdin:
            line = line.strip()
            if not line:
                continue
            try:
                reply = {"ok": True, "reply": run(lib, [str(a) for a in json.loads(line)])}
            except BaseException as exc:   # noqa: BLE001 -- the client turns it into a test failure
                reply = {"ok": False, "error": repr(exc)}
            sys.stdout.write(json.dumps(reply) + "\n")
            sys.stdout.flush()
        return
    reply = run(lib, sys.argv[1:])
    if reply is not None:
        print(json.dumps(reply))


def apply_options(lib, coder, opts, decoder):
    fn = lib.gpujpeg_decoder_set_option if decoder else lib.gpujpeg_encoder_set_option
    fn.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    for o in opts:
        key, _, val = o.partition("=")
        assert fn(coder, key.encode(), val.encode()) == 0, o


def run(lib, a):
    """one command: a = [mode, arguments...] as on the command line; returns a JSON-able reply or None.
    Arguments of the form opt:<key>=<value> are passed to gpujpeg_{en,de}coder_set_option of the coder the command creates."""
    opts = [x[4:] for x in a if x.startswith("opt:")]
    pars = dict(x[4:].split("=") for x in a if x.startswith("par:"))   # par:<field>=<int>: struct gpujpeg_parameters fields
    subs = [int(x[4:], 0) for x in a if x.startswith("sub:")]           # sub:<packed>: gpujpeg_parameters_chroma_subsampling
    a = [x for x in a if not x.startswith(("opt:", "par:", "sub:"))]
    mode = a[0]
    if mode == "encode":
        kind, w, h, q, rst, il, path = a[1], *map(int, a[2:7]), a[7]
        img = gen(kind, w, h)
        enc = lib.gpujpeg_encoder_create(None)
        p, pi = params(lib, w, h, q, rst, il)
        for k, v in pars.items():
            setattr(p, k, int(v))
        if len(a) > 9:   # chroma subsampling: GPUJPEG_SUBSAMPLING_xxx packing, first component on top
            lib.gpujpeg_parameters_chroma_subsampling.argtypes = [C.POINTER(Param), C.c_uint32]
            lib.gpujpeg_parameters_chroma_subsampling(C.byref(p), int(a[8]) << 28 | int(a[9]) << 24 | 0x111100)
        encode(lib, enc, img, p, pi).tofile(path)
        lib.gpujpeg_encoder_destroy(enc)
    elif mode == "encode_raw":
        src, fmt, cs, w, h, q, rst, il, path = a[1], *map(int, a[2:9]), a[9]
        raw = np.fromfile(src, np.uint8)
        enc = lib.gpujpeg_encoder_create(None)
        apply_options(lib, enc, opts, False)
        p, pi = params(lib, w, h, q, rst, il)
        pi.pixel_format, pi.color_space = fmt, cs   # comp_count stays 0: sampling follows the pixel format
        if len(a) > 10:
            p.color_space_internal = int(a[10])   # e.g. 1 = GPUJPEG_RGB: RGB-internal JPEG (Adobe APP14)
        for packed in subs:   # e.g. 0x11111111: four components (an RGBA image keeps its alpha)
            lib.gpujpeg_parameters_chroma_subsampling.argtypes = [C.POINTER(Param), C.c_uint32]
            lib.gpujpeg_parameters_chroma_subsampling(C.byref(p), packed)
        for k, v in pars.items():
            setattr(p, k, int(v))
        encode(lib, enc, raw, p, pi).tofile(path)
        lib.gpujpeg_encoder_destroy(enc)
    elif mode == "decode_fmt":
        data = np.fromfile(a[1], np.uint8)
        dec = lib.gpujpeg_decoder_create(None)
        apply_options(lib, dec, opts, True)
        lib.gpujpeg_decoder_set_output_format(dec, int(a[2]), int(a[3]))
        out = DecOut()
        out.type = 0
        assert lib.gpujpeg_decoder_decode(dec, data.ctypes.data, data.size, C.byref(out)) == 0
        np.ctypeslib.as_array((C.c_uint8 * out.data_size).from_address(out.data)).tofile(a[4])
        reply = {"pixel_format": out.param_image.pixel_format, "color_space": out.param_image.color_space,
                 "size": out.data_size}
        if out.metadata:   # struct gpujpeg_image_metadata: {rotation:2, flip:1} bit-field word, then the "set" bit
            m = (C.c_uint32 * 2).from_address(out.metadata)
            reply["orientation"] = [int(m[1] & 1), int(m[0] & 3), int((m[0] >> 2) & 1)]
        lib.gpujpeg_decoder_destroy(dec)
        return reply
    elif mode == "file_props":   # gpujpeg_image_get_properties of the reference: [rc, width, height, colour space, pixel format]
        pi = ImgParam()
        pi.pixel_format = -1
        lib.gpujpeg_image_get_properties.argtypes = [C.c_char_p, C.POINTER(ImgParam), C.c_int]
        rc = lib.gpujpeg_image_get_properties(a[1].encode(), C.byref(pi), int(a[2]))
        return [rc, pi.width, pi.height, pi.color_space, pi.pixel_format]
    elif mode == "file_load":    # gpujpeg_image_load_from_file of the reference -> raw dump
        lib.gpujpeg_image_load_from_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        ptr, size = C.c_void_p(), C.c_size_t(0)
        rc = lib.gpujpeg_image_load_from_file(a[1].encode(), C.byref(ptr), C.byref(size))
        if rc == 0:
            np.ctypeslib.as_array((C.c_uint8 * size.value).from_address(ptr.value)).tofile(a[2])
        return [rc, size.value]
    elif mode == "file_save":    # gpujpeg_image_save_to_file of the reference: raw dump -> <out> as <fmt> <cs> <w> <h>
        raw = np.fromfile(a[1], np.uint8)
        pi = ImgParam()
        lib.gpujpeg_image_set_default_parameters(C.byref(pi))
        pi.pixel_format, pi.color_space, pi.width, pi.height = int(a[3]), int(a[4]), int(a[5]), int(a[6])
        lib.gpujpeg_image_save_to_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(ImgParam)]
        name = C.create_string_buffer(a[2].encode())
        return [lib.gpujpeg_image_save_to_file(name, raw.ctypes.data, raw.size, C.byref(pi)), name.value.decode()]
    elif mode == "decode":
        data = np.fromfile(a[1], np.uint8)
        dec = lib.gpujpeg_decoder_create(None)
        lib.gpujpeg_decoder_set_output_format(dec, 1, 1)  # GPUJPEG_RGB, GPUJPEG_444_U8_P012
        out = DecOut()
        out.type = 0
        assert lib.gpujpeg_decoder_decode(dec, data.ctypes.data, data.size, C.byref(out)) == 0
        np.ctypeslib.as_array((C.c_uint8 * out.data_size).from_address(out.data)).tofile(a[2])
        lib.gpujpeg_decoder_destroy(dec)
    elif mode == "bench":
        kind, w, h, q, rst, iters = a[1], *map(int, a[2:7])
        img, pinned, _keep = pinned_copy(lib, gen(kind, w, h))
        enc = lib.gpujpeg_encoder_create(None)
        # the decoder takes perf_stats at create time (gpujpeg_decoder_init_parameters); with plain
        # gpujpeg_decoder_create its timers stay off and get_stats reports nothing (round 1: decode_ms_gpu 0.0)
        class DecInit(C.Structure):
            _fields_ = [("stream", C.c_void_p), ("verbose", C.c_int), ("perf_stats", C.c_bool), ("ff_cs_itu601_is_709", C.c_bool)]
        lib.gpujpeg_decoder_create_with_params.restype = C.c_void_p
        lib.gpujpeg_decoder_create_with_params.argtypes = [C.POINTER(DecInit)]
        di = DecInit(None, -1, True, False)
        dec = lib.gpujpeg_decoder_create_with_params(C.byref(di))
        lib.gpujpeg_decoder_set_output_format(dec, 1, 1)
        p, pi = params(lib, w, h, q, rst, 0, perf=1)
        jpeg, _, _keep2 = pinned_copy(lib, encode(lib, enc, img, p, pi))
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
        lib.gpujpeg_encoder_destroy(enc)
        lib.gpujpeg_decoder_destroy(dec)
        return {"impl": "reference-gpu", "w": w, "h": h, "kind": kind, "jpeg_bytes": int(jpeg.size),
                "encode_ms_e2e": med(te), "encode_ms_gpu": med(tg), "decode_ms_e2e": med(td),
                "decode_ms_gpu": med(tdg), "iters": iters, "input": "pinned" if pinned else "pageable",
                "method": "serial gpujpeg_encoder_encode / gpujpeg_decoder_decode calls, host buffers, median of %d "
                          "(the reference's own `gpujpegtool -n` loop, src/main.c:805-815)" % iters}
    return None


if __name__ == "__main__":
    main()
