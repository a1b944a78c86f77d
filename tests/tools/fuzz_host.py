"""Mutation fuzzing of the host code that reads untrusted bytes: the marker reader incl. SPIFF directory and Exif (gj_codestream.c,
gj_exif.c), the enc_exif_tag option parser, and the PNM / PAM / Y4M header parsers (gj_imageio.c).  Meant to run under the sanitizers:

    gcc -O1 -g -fsanitize=address,undefined -shared -fPIC ... (the two shims of tests/_shims.py) over tests/cpu_shims/{host,io}_shim.so
    LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 python tests/tools/fuzz_host.py

Round 2: 60 000 + 40 000 + 30 000 cases, no report.  tests/test_fuzz_host.py runs a short version without sanitizers (no crash, no hang)."""
import ctypes as C, os, sys, random, tempfile
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import _oracle as o
from _shims import hs, io
hs.shim_parse_meta.argtypes = [np.ctypeslib.ndpointer(np.uint8), C.c_size_t, np.ctypeslib.ndpointer(np.int32)]
hs.shim_add_exif_tag.argtypes = [C.c_char_p]
rng = random.Random(7)
# seeds: real headers of every flavour
seeds = []
img = o.gen_image("photo", 64, 48)
for q, rst, il in ((75, 4, 0), (90, 0, 1)):
    seeds.append(bytes(o.encode(img, q, rst, il)))
for ht in (2, 8):
    hs.shim_set_header_type(ht); hs.shim_set_orientation(1, 3, 1)
    hs.shim_add_exif_tag(b"0x10F:ASCII=maker with a long name to have offsets"); 
    out = np.zeros(70000, np.uint8); n = hs.shim_header2(64, 48, 75, 4, 0, 3, 1, 1, 3, out)
    seeds.append(bytes(out[:n]) + seeds[0][600:])
    hs.shim_clear_exif_tags(); hs.shim_set_header_type(0); hs.shim_set_orientation(0, 0, 0)
rgba = np.zeros((16, 16, 4), np.uint8)
seeds.append(bytes(o.encode_any(rgba, 16, 16, 6, 1, 75, 2, 1, (1, 1), alpha=True)))
n_ok = 0
res = np.zeros(5, np.int32)
N = int(os.environ.get('GJ_FUZZ_N', '60000'))
for it in range(N):
    s = bytearray(rng.choice(seeds))
    lim = min(len(s), 900)
    for _ in range(rng.randint(1, 6)):
        k = rng.random()
        p = rng.randrange(min(lim, len(s)) or 1) if len(s) else 0
        if not s: break
        if k < 0.5: s[p] = rng.randrange(256)
        elif k < 0.7: s[p] = rng.choice([0, 0xFF, 0x7F, 0x80, 1])
        elif k < 0.85: del s[p:p + rng.randint(1, 8)]
        else: s[p:p] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 8)))
    if rng.random() < 0.2: s = s[:rng.randrange(4, lim)]
    a = np.frombuffer(bytes(s), np.uint8).copy()
    n_ok += hs.shim_parse_meta(a, a.size, res) == 0
print("reader fuzz done", n_ok, "parsed")
# exif tag option strings
alphabet = "0123456789xXabcdef:=,/- ASCIHORTLNGBYEUDFWPmk."
for it in range(N * 2 // 3):
    t = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 40)))
    if rng.random() < 0.5: t = rng.choice(["0x112:SHORT=", "Orientation=", "XResolution=", "0x9286:UNDEFINED=", "0x13E:RATIONAL=", "0x100:LONG="]) + t
    hs.shim_add_exif_tag(t.encode())
    if it % 50 == 0:
        out = np.zeros(70000, np.uint8); hs.shim_set_header_type(8); hs.shim_header2(64, 48, 75, 4, 0, 3, 1, 1, 3, out); hs.shim_set_header_type(0)
        hs.shim_clear_exif_tags()
print("exif option fuzz done")
# image files
class IP(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("cs", C.c_int), ("pf", C.c_int), ("pad", C.c_int)]
io.gpujpeg_image_get_properties.argtypes = [C.c_char_p, C.POINTER(IP), C.c_int]
io.gpujpeg_image_load_from_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
io.gpujpeg_image_destroy.argtypes = [C.c_void_p]
fseeds = [b"P6\n# c\n5 4\n255\n" + bytes(60), b"P5\n3 3\n255\n" + bytes(9), b"P7\nWIDTH 2\nHEIGHT 2\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n" + bytes(16),
          b"YUV4MPEG2 W6 H4 F25:1 Ip A0:0 C420jpeg XCOLORRANGE=FULL\nFRAME\n" + bytes(36), b"YUV4MPEG2 W4 H2 Cmono\nFRAME\n" + bytes(8)]
d = tempfile.mkdtemp()
devnull = os.open(os.devnull, os.O_WRONLY); os.dup2(devnull, 2)
for it in range(N // 2):
    s = bytearray(rng.choice(fseeds))
    for _ in range(rng.randint(1, 5)):
        p = rng.randrange(min(len(s), 70))
        k = rng.random()
        if k < 0.6: s[p] = rng.choice(b"0123456789 \n#PYWHC-" ) if rng.random() < 0.7 else rng.randrange(256)
        elif k < 0.8: del s[p:p + rng.randint(1, 6)]
        else: s[p:p] = bytes(rng.choice(b"0123456789 \n") for _ in range(rng.randint(1, 6)))
    ext = rng.choice(["pnm", "pam", "y4m", "ppm", "pgm"])
    fn = os.path.join(d, "f." + ext)
    open(fn, "wb").write(bytes(s))
    ip = IP()
    io.gpujpeg_image_get_properties(fn.encode(), C.byref(ip), 1)
    ptr, size = C.c_void_p(), C.c_size_t(0)
    if io.gpujpeg_image_load_from_file(fn.encode(), C.byref(ptr), C.byref(size)) == 0:
        io.gpujpeg_image_destroy(ptr)
os.write(1, b"file fuzz done\n")
