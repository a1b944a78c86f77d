"""Short run of tests/tools/fuzz_host.py (mutated streams, option strings and image-file headers through the host code): must end
without a crash or a hang.  The long run under AddressSanitizer / UBSan is described in the tool."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_host_parsers_survive_mutated_input():
    env = dict(os.environ, GJ_FUZZ_N="3000")
    r = subprocess.run([sys.executable, os.path.join(HERE, "tools", "fuzz_host.py")], env=env, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert b"reader fuzz done" in r.stdout and b"exif option fuzz done" in r.stdout and b"file fuzz done" in r.stdout
