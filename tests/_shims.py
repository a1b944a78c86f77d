"""Builds and loads the two CPU shims under tests/cpu_shims (test infrastructure):
   kernel_math.so : the kernels' per-thread arithmetic (gj_device.cuh) compiled for the host
   host_shim.so   : internal host functions of the product (tables, writer, reader)
   io_shim.so     : the product's public image-file helpers (gj_common.c + gj_imageio.c) over host stand-ins for CUDA"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SH = os.path.join(HERE, "cpu_shims")
CSRC = os.path.join(ROOT, "gpujpeg_b200", "csrc")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def _build():
    km = os.path.join(SH, "kernel_math.so")
    if _stale(km, [os.path.join(SH, "kernel_math.cpp"), os.path.join(CSRC, "gj_device.cuh")]):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", km,
                               os.path.join(SH, "kernel_math.cpp")])
    hs = os.path.join(SH, "host_shim.so")
    srcs = [os.path.join(SH, "host_shim.c"), os.path.join(CSRC, "gj_tables.c"), os.path.join(CSRC, "gj_codestream.c"),
            os.path.join(CSRC, "gj_exif.c"), os.path.join(SH, "names_stub.c")]
    if _stale(hs, srcs + [os.path.join(CSRC, "gj_internal.h")]):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-std=gnu11", "-shared", "-fPIC", "-o", hs] + srcs)
    io = os.path.join(SH, "io_shim.so")
    srcs = [os.path.join(SH, "cuda_stub.c")] + [os.path.join(CSRC, f) for f in ("gj_common.c", "gj_imageio.c", "gj_tables.c",
                                                                                "gj_codestream.c", "gj_exif.c")]
    if _stale(io, srcs + [os.path.join(CSRC, "gj_internal.h")]):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-std=gnu11", "-shared", "-fPIC", "-o", io] + srcs)
    return km, hs, io


_km, _hs, _io = _build()
io = C.CDLL(_io)
km = C.CDLL(_km)
km.km_check_rgb_to_ycbcr_exhaustive.restype = C.c_long
km.km_check_ycbcr_to_rgb_exhaustive.restype = C.c_long
km.km_fdct_quant_plane.argtypes = [_u8p, C.c_int, C.c_int, np.ctypeslib.ndpointer(np.float32), _i16p]
km.km_idct_plane.argtypes = [_i16p, C.c_int, C.c_int, np.ctypeslib.ndpointer(np.uint16), C.c_int, _u8p]
km.km_value_bits.restype = C.c_uint
hs = C.CDLL(_hs)
hs.shim_header.argtypes = [C.c_int] * 5 + [_u8p]
hs.shim_forward_table_zz.argtypes = [C.c_int, C.c_int, np.ctypeslib.ndpointer(np.float32), _u8p]
hs.shim_enc_lut.argtypes = [C.c_int, np.ctypeslib.ndpointer(np.uint32), np.ctypeslib.ndpointer(np.uint32)]
hs.shim_dec_lut_symbol.argtypes = [C.c_int, C.c_int, C.c_uint, C.POINTER(C.c_int)]
hs.shim_dec_fast_symbol.argtypes = [C.c_int, C.c_int, C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
hs.shim_parse.argtypes = [_u8p, C.c_size_t, np.ctypeslib.ndpointer(np.int32), np.ctypeslib.ndpointer(np.uint32),
                          np.ctypeslib.ndpointer(np.uint32), C.c_int]
hs.shim_header2.argtypes = [C.c_int] * 9 + [np.ctypeslib.ndpointer(np.uint8)]
hs.shim_raw_layout.argtypes = [C.c_int] * 4 + [np.ctypeslib.ndpointer(np.int64)]
hs.shim_geometry.argtypes = [C.c_int] * 4 + [np.ctypeslib.ndpointer(np.int64)]
hs.shim_geometry_ss.argtypes = [C.c_int] * 6 + [np.ctypeslib.ndpointer(np.int64)]
