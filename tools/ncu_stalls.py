#!/usr/bin/env python
"""Where do warps wait?  Aggregates ncu's warp-stall samples per source line (and the barrier each BAR.SYNC belongs to).

    python tools/ncu_stalls.py gpurun_out/x.ncu-rep k_encode_super [N]
"""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
by_line = collections.Counter(); by_reason = collections.Counter(); line_reason = collections.defaultdict(collections.Counter)
cur = "?"
total = 0
for r in rows:
    if len(r) > 3 and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    n = int(d["# Samples"] or 0)
    src = d["Source"]
    key = src.strip()[:110]
    by_line[key] += n; total += n
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
            by_reason[k] += int(v); line_reason[key][k] += int(v)
print("total samples", total)
print("by reason:", ", ".join(f"{k[6:]}={v * 100.0 / max(total, 1):.1f}%" for k, v in by_reason.most_common(10)))
for k, v in by_line.most_common(top):
    rs = ", ".join(f"{a[6:]}={b}" for a, b in line_reason[k].most_common(3))
    print(f"{v * 100.0 / max(total, 1):5.1f}%  {k}   [{rs}]")
