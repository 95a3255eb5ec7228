#!/usr/bin/env python
"""Static SASS evidence for profiles/: per kernel of libt3c.so, how many instructions and how many of the opcodes that show which
memory path a kernel uses (bulk / tensor async copies and their barriers, 128-bit global accesses, shared-memory traffic).

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ternary_image_codec_b200", "libt3c.so")
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDG.E.128", "STG.E.128", "LDG.E.64", "STG.E.64", "LDS", "STS", "LOP3", "IMAD", "IDP", "PRMT", "REDUX", "ATOMG", "REDG", "LDL", "STL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    print("# cuobjdump -sass ternary_image_codec_b200/libt3c.so (sm_100a), static counts per kernel: total instructions, then the watched opcodes that occur")
    print("# UBLKCP = cp.async.bulk, UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA tile load / store), SYNCS = mbarrier operations, LDL/STL = local-memory (spill) traffic")
    for name, c in sorted(per.items(), key=lambda kv: demangle(kv[0])):
        d = demangle(name)
        d = re.sub(r"\(anonymous namespace\)::|<unnamed>::|t3c::|\((?:int|bool|unsigned int)\)", "", d).replace("void ", "")
        d = d[:d.index("(")] if "(" in d else d
        print(f"{d:48s} total {c['total']:6d}  " + "  ".join(f"{w} {c[w]}" for w in WATCH if c[w]))


if __name__ == "__main__":
    main()
