"""Correlates an ncu report's per-SASS-instruction samples with CUDA source lines (via nvdisasm -g line info of the
cubin inside the built library) and prints the hottest source lines.
    python profiles/srcsum.py gpurun_out/prof.ncu-rep k_huff_decode gj_huffman [top_n]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern, cubin_stem = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gpujpeg_b200", "lib", "libgpujpeg.so.0")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith(cubin_stem) and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# walk disassembly: track current function and current source line
lines_of = {}
func, cur = None, None
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        func = m.group(1)
        lines_of.setdefault(func, [])
        cur = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and func:
        lines_of[func].append((int(m.group(1), 16), cur, m.group(2)))
fn = [f for f in lines_of if (sys.argv[5] if len(sys.argv) > 5 else kern) in f][0]
ins = lines_of[fn]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
rd = list(csv.reader(out[start:]))
hdr = rd[0]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
rows = [r for r in rd[1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
base = int(rows[0][0], 16)
agg = {}
tot_s = tot_i = 0.0
for r in rows:
    off = int(r[0], 16) - base
    k = min(range(len(ins)), key=lambda j: abs(ins[j][0] - off)) if off >= len(ins) * 16 else off // 16
    k = min(k, len(ins) - 1)
    line = ins[k][1]
    s, n = float(r[idx["# Samples"]] or 0), float(r[idx["Instructions Executed"]] or 0)
    a = agg.setdefault(line, [0.0, 0.0, {}])
    a[0] += s
    a[1] += n
    for st in stalls:
        v = float(r[idx[st]] or 0)
        if v:
            a[2][st[6:]] = a[2].get(st[6:], 0) + v
    tot_s += s
    tot_i += n
srcs = {}
print("kernel %s: %d SASS instructions, %.0f warp-instructions executed, %.0f samples" % (kern, len(rows), tot_i, tot_s))
for line, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    text = ""
    if line:
        path = os.path.join(ROOT, "gpujpeg_b200", "csrc", line[0])
        if path not in srcs and os.path.exists(path):
            srcs[path] = open(path).read().splitlines()
        if path in srcs and line[1] <= len(srcs[path]):
            text = srcs[path][line[1] - 1].strip()[:80]
    top = " ".join("%s=%.0f%%" % (k, v / s * 100) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3]) if s else ""
    print("%5.1f%% smp %5.1f%% ins  %-22s %-80s %s" % (s / tot_s * 100, n / tot_i * 100, "%s:%d" % line if line else "?", text, top))
