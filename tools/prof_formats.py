"""Short driver for ncu: the streaming kernels either side of the path on 8K-sized inputs (no timing).
    python tools/prof_formats.py bridge|base243|t3v"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import secondary as S2

which = sys.argv[1] if len(sys.argv) > 1 else "bridge"
ctx = S2.Ctx(0)
r = {"bridge": S2.bridge8k, "base243": S2.formats8k, "t3v": S2.t3v8k}[which](ctx)
print({k: v for k, v in r.items() if k.endswith("_us")})
