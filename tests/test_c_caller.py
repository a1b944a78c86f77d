"""A plain C application written against <libgpujpeg/gpujpeg.h> (tests/c_api/roundtrip.c), compiled with gcc against
include/ and linked with -lgpujpeg: source- and link-compatibility of the drop-in boundary.
CPU: it builds, links and reports "no device"; GPU: its files equal the oracle's."""
import os
import subprocess

import numpy as np
import pytest

import _oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build(tmp_path):
    from gpujpeg_b200 import build as b
    lib = b.build_library()
    exe = str(tmp_path / "roundtrip")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", os.path.join(HERE, "c_api", "roundtrip.c"), "-I",
                           os.path.join(ROOT, "include"), "-L", os.path.dirname(lib), "-lgpujpeg", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(lib) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return exe, env


def test_c_application_builds_links_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    exe, env = build(tmp_path)
    rc = subprocess.call([exe, "64", "48", str(tmp_path / "a.jpg"), str(tmp_path / "a.rgb")], env=env,
                         stderr=subprocess.DEVNULL)
    assert rc == 3, "without a CUDA device the library must refuse, not fall back to a CPU path"


@pytest.mark.gpu
def test_c_application_round_trip_equals_oracle(tmp_path):
    exe, env = build(tmp_path)
    w, h = 322, 200
    out = subprocess.check_output([exe, str(w), str(h), str(tmp_path / "a.jpg"), str(tmp_path / "a.rgb")], env=env)
    assert out.decode().startswith("322x200 4:2:0")
    img = np.repeat((np.arange(h) * 255 // h).astype(np.uint8)[:, None, None], w, axis=1).repeat(3, axis=2)
    want = o.encode(np.ascontiguousarray(img), 80, 4, 1, sampling=(2, 2))
    got = np.fromfile(tmp_path / "a.jpg", np.uint8)
    assert np.array_equal(got, want)
    assert np.array_equal(np.fromfile(tmp_path / "a.rgb", np.uint8).reshape(h, w, 3), o.decode(want))
