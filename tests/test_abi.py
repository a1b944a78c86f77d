"""The C-ABI library loads on a GPU-less machine, exports every function include/gpujpeg_b200.h declares,
keeps the reference's struct layout, and fails loudly (NULL / -1) instead of falling back.  CPU only."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gpujpeg_b200.h")


EXT_HEADER = os.path.join(ROOT, "include", "gpujpegx.h")


def declared_functions(header=HEADER):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"GPUJPEG_API[^;{]*?\b(gpujpegx?_\w+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    import gpujpeg_b200
    names = declared_functions()
    assert len(names) >= 70
    missing = [n for n in names if not hasattr(gpujpeg_b200.lib, n)]
    assert not missing, missing
    # the additive extension header (resident re-runs, coefficient read-back, multi-GPU batches)
    ext = declared_functions(EXT_HEADER)
    assert len(ext) >= 11 and all(n.startswith("gpujpegx_") for n in ext)
    missing = [n for n in ext if not hasattr(gpujpeg_b200.lib, n)]
    assert not missing, missing


def test_struct_layout_matches_reference_abi(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include <libgpujpeg/gpujpeg.h>\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(struct gpujpeg_parameters),sizeof(struct gpujpeg_image_parameters),'
                   'sizeof(struct gpujpeg_encoder_input),sizeof(struct gpujpeg_decoder_output),'
                   'sizeof(struct gpujpeg_decoder_init_parameters),sizeof(struct gpujpeg_image_info),'
                   'sizeof(struct gpujpeg_duration_stats),sizeof(struct gpujpeg_devices_info),'
                   'sizeof(struct gpujpeg_image_metadata),offsetof(struct gpujpeg_parameters,quality),'
                   'offsetof(struct gpujpeg_parameters,sampling_factor),'
                   'offsetof(struct gpujpeg_parameters,color_space_internal));return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    # sizes/offsets probed from the reference headers
    assert [int(x) for x in out] == [40, 20, 24, 64, 16, 512, 72, 3048, 8, 8, 28, 36]


def test_version_and_pure_host_api():
    import gpujpeg_b200 as g
    assert g.version() == "0.27.13"
    lib = g.lib
    lib.gpujpeg_subsampling_get_name.restype = C.c_char_p
    p = g.api.Parameters()
    lib.gpujpeg_set_default_parameters(C.byref(p))
    assert (p.quality, p.restart_interval, p.interleaved, p.color_space_internal) == (75, 8, 0, 3)
    # strings pinned by the reference's unit test (test/unit/run_tests.c:17-36)
    for sub, name in ((0x11111100, b"4:4:4"), (0x21111100, b"4:2:2"), (0x22111100, b"4:2:0"), (0x41111100, b"4:1:1"),
                      (0x11111111, b"4:4:4:4"), (0x21111121, b"4:2:2:4")):
        lib.gpujpeg_parameters_chroma_subsampling(C.byref(p), C.c_uint32(sub))
        assert lib.gpujpeg_subsampling_get_name(p.comp_count, p.sampling_factor) == name
    pi = g.api.image_parameters(1920, 1080)
    assert lib.gpujpeg_image_calculate_size(C.byref(pi)) == 1920 * 1080 * 3
    assert lib.gpujpeg_encoder_suggest_restart_interval(C.byref(g.api.image_parameters(7680, 4320)), 0x11111100, False, 0) == 36
    assert lib.gpujpeg_encoder_suggest_restart_interval(C.byref(pi), 0x11111100, False, 0) == 24


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gpujpeg_b200 as g
    with pytest.raises(g.GpuJpegError):
        g.Encoder()
    with pytest.raises(g.GpuJpegError):
        g.Decoder()


def test_product_does_not_reference_oracle():
    # the product tree must never import, link or call anything under oracle/
    bad = []
    pkg = os.path.join(ROOT, "gpujpeg_b200")
    for dp, _, files in os.walk(pkg):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".c", ".cu", ".cuh", ".h", ".py")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"#include\s+\"[^\"]*oracle|import\s+_?oracle|liboracle|oracle/", txt):
                    bad.append(f)
    assert not bad, bad
