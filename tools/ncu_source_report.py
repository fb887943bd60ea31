"""Per-source-line view of one ncu capture of an encode kernel.

    python tools/ncu_source_report.py <capture.ncu-rep> <kernel-mangled-substring> [top N]

Joins the SASS rows of `ncu --page source --csv` (executed instructions, stall samples, shared-memory wavefronts per
instruction) with the file:line of every instruction from `nvdisasm --print-line-info` on the cubin inside
zig-flac_b200/libzigflac_b200.so (the library must be the build that was profiled; compile with -lineinfo), and prints
totals per source line: where the instructions, the stall samples and the excess shared-memory wavefronts (bank
conflicts) are.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "zig-flac_b200", "libzigflac_b200.so")], cwd=tmp,
               capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("zf_decode" if "zf_dec" in kern or "3dec" in kern else "zf_capi")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = dis.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
line_of = []  # per instruction, in order
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        line_of.append(cur)

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
assert abs(len(body) - len(line_of)) <= 64, (len(body), len(line_of))


def num(r, name):
    try:
        return float(r[col[name]])
    except Exception:
        return 0.0


agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for i, r in enumerate(body):
    key = line_of[i] if i < len(line_of) else ("?", 0)
    a = agg[key]
    vals = {"inst": num(r, "Instructions Executed"), "samples": num(r, "# Samples"),
            "wf": num(r, "L1 Wavefronts Shared"), "wf_excess": num(r, "L1 Wavefronts Shared Excessive")}
    for h in stall_cols:
        vals[h] = num(r, h)
    for k, v in vals.items():
        a[k] += v
        tot[k] += v

src = {}


def text(key):
    f, n = key
    if f not in src:
        p = os.path.join(ROOT, "zig-flac_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""


print(f"total: {tot['inst']:.0f} warp instructions, {tot['samples']:.0f} samples, shared wavefronts {tot['wf']:.0f} "
      f"(excess {tot['wf_excess']:.0f})")
print("stall samples:", ", ".join(f"{h[6:]} {tot[h] / max(tot['samples'], 1) * 100:.1f}%" for h in
                                  sorted(stall_cols, key=lambda h: -tot[h])[:8]))
for title, k in (("instructions", "inst"), ("stall samples", "samples"), ("excess shared wavefronts", "wf_excess")):
    print(f"\n== top lines by {title}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][k])[:top]:
        if a[k] <= 0:
            break
        big = max(stall_cols, key=lambda h: a[h])
        print(f"{a[k] / max(tot[k], 1) * 100:6.2f}%  inst {a['inst'] / max(tot['inst'], 1) * 100:5.2f}%  smp {a['samples'] / max(tot['samples'], 1) * 100:5.2f}%"
              f"  wf+ {a['wf_excess']:10.0f}  {big[6:]:<12} {key[0]}:{key[1]:<5} {text(key)}")
