"""Join an ncu source-page export with nvdisasm line info: samples / executed warp instructions per source line.
    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all armada_b200/libarmada_b200.so; nvdisasm -g -c *.cubin > all.sass
    python tools/ncu_lines.py src.csv all.sass 'k_schedule_passILi3E' [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, fn = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line_of = {}
cur, on = None, False
for ln in open(sass, errors="replace"):
    if ln.startswith(".text."):
        on = fn in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = {n: i for i, n in enumerate(rows[hdr])}
base = None
samples, execd = defaultdict(int), defaultdict(int)
stall = defaultdict(lambda: defaultdict(int))
tot_s = tot_e = 0
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    a = int(r[H["Address"]], 16)
    if base is None:
        base = a
    key = line_of.get(a - base, ("?", 0))
    s, e = int(r[H["# Samples"]] or 0), int(r[H["Instructions Executed"]] or 0)
    samples[key] += s
    execd[key] += e
    tot_s += s
    tot_e += e
    for n in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_lg", "stall_sleep", "stall_barrier", "stall_membar", "stall_branch_resolving"):
        if n in H:
            stall[key][n] += int(r[H[n]] or 0)
print(f"total samples {tot_s}, warp instructions {tot_e}")
src_cache = {}
for key, s in sorted(samples.items(), key=lambda kv: -kv[1])[:top]:
    f, l = key
    text = ""
    try:
        if f not in src_cache:
            import glob
            p = glob.glob(f"armada_b200/csrc/{f}") + glob.glob(f"/usr/local/cuda/include/**/{f}", recursive=True)
            src_cache[f] = open(p[0]).read().split("\n") if p else []
        text = src_cache[f][l - 1].strip()[:110] if src_cache[f] else ""
    except Exception:
        pass
    st = ", ".join(f"{k[6:]} {v}" for k, v in sorted(stall[key].items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100.0 * s / max(tot_s, 1):5.1f}%  {s:8d} smp {execd[key]:10d} inst  {f}:{l:<5d} [{st}] | {text}")
