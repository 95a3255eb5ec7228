"""Secondary workloads of BASELINE.json (configs[2..4]) -- one JSON line per workload (tools/secondary.py does the work; bench.py
carries the same numbers in its `secondary` object)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import secondary as S2

if __name__ == "__main__":
    ctx = S2.Ctx(0)
    def words():
        return [S2.words8k(ctx, 2, "words8k_k20: encode_profile_from_raw + consistent decode on 16.6 M raw words, RS(26,20) 1D"),
                S2.words8k(ctx, 1, "words8k_default: the reference's default EncoderContext (P2, uniform k=22), raw words in/out")]
    which = sys.argv[1:] or ["words8k", "raw8k", "uep2d", "t3v8k", "formats8k", "stream240"]
    for w in which:
        r = {"words8k": words, "raw8k": lambda: S2.raw8k(ctx), "uep2d": lambda: S2.uep2d(ctx), "t3v8k": lambda: S2.t3v8k(ctx),
             "formats8k": lambda: S2.formats8k(ctx), "stream240": lambda: S2.stream240(ctx)}[w]()
        for x in (r if isinstance(r, list) else [r]):
            print(json.dumps(x))
