"""Thread scaling of the CPU oracle on this host (picks the thread count bench.py's reference arm uses)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as o
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("no cgroup cpu.max", e)
w, h = 7680, 4320
img = o.gen_image("photo", w, h)
for th in (1, 8, 16, 32, 64, 128):
    if th > (os.cpu_count() or 1):
        break
    j = o.encode(img, 75, 36, threads=th)
    t = time.time(); j = o.encode(img, 75, 36, threads=th); te = time.time() - t
    t = time.time(); d = o.decode(j, threads=th); td = time.time() - t
    print("%3d threads: enc %.3f s dec %.3f s -> %.1f Mpix/s" % (th, te, td, w * h / 1e6 / (te + td)))
