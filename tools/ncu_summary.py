#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/r1_prof_fast.ncu-rep profiles/r01_fast_summary.txt

Per kernel: duration, DRAM bytes, pipe / issue / LSU figures, stall mix, SASS opcode histogram and the source
lines (needs -lineinfo + --import-source on) that execute the most instructions / shared-memory wavefronts.
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    w = io.StringIO()
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    names = []
    done_raw = set()
    for r in raw[2:]:
        name = r[hdr.index("Kernel Name")]
        if name in done_raw:
            continue
        done_raw.add(name)
        short = name.split("(")[0].split("::")[-1]
        names.append(short)
        w.write(f"== {short}  [{name[:110]}]\n")
        for k in RAW_KEYS:
            if k in hdr:
                w.write(f"   {k:72s} {r[hdr.index(k)]} {units[hdr.index(k)]}\n")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        w.write("   stall mix (warps stalled per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:9]) + "\n")
    seen = set()
    for short in names:
        if short in seen:
            continue
        seen.add(short)
        txt = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + short.split("<")[0]])
        rows = list(csv.reader(io.StringIO(txt)))
        hs = [r for r in rows if "Instructions Executed" in r]
        if not hs:
            continue
        h = hs[0]
        iline, isrc, iaddr, isass = 0, 1, 2, 3
        col = lambda name: h.index(name) if name in h else -1           # kernels without shared memory have no wavefront columns
        ie, iw, ii = h.index("Instructions Executed"), col("L1 Wavefronts Shared"), col("L1 Wavefronts Shared Ideal")
        ismp = h.index("# Samples") if "# Samples" in h else -1
        stall_cols = [(i, c[6:]) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        samples = collections.OrderedDict()   # (file, line) -> [n, Counter(reason)]
        fpath = ""
        lines = collections.OrderedDict()
        ops = collections.Counter()
        first_launch_done = False
        for r in rows:
            if len(r) == 2 and r[0] == "File Path":
                fpath = r[1].split("/")[-1]
            if len(r) == 2 and r[0] == "Function Name" and lines and fpath == "":
                pass
            if len(r) <= ie:
                continue
            if r[iline].isdigit() and r[ie].isdigit():
                key = (fpath, int(r[iline]))
                e = lines.setdefault(key, [0, 0, 0, r[isrc].strip()[:100]])
                e[0] += int(r[ie]); e[1] += int(r[iw]) if iw >= 0 and r[iw].isdigit() else 0; e[2] += int(r[ii]) if ii >= 0 and r[ii].isdigit() else 0
                if ismp >= 0 and r[ismp].isdigit():
                    se = samples.setdefault(key, [0, collections.Counter(), r[isrc].strip()[:100]])
                    se[0] += int(r[ismp])
                    for i, name in stall_cols:
                        if r[i].isdigit() and r[i] != "0":
                            se[1][name] += int(r[i])
            elif r[iline] == "" and r[ie].isdigit():
                op = r[isass].split()
                if op:
                    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
                    ops[o.split(".")[0]] += int(r[ie])
        tot = sum(v[0] for v in lines.values()) or 1
        totw = sum(v[1] for v in lines.values()) or 1
        w.write(f"\n== {short}: source-line profile (all profiled launches of this kernel summed; {tot} warp instructions, {totw} smem wavefronts)\n")
        w.write("   opcode mix: " + ", ".join(f"{k} {100 * v / max(1, sum(ops.values())):.1f}%" for k, v in ops.most_common(18)) + "\n")
        w.write("   top lines by warp instructions executed:\n")
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:28]:
            w.write(f"     {100 * v[0] / tot:5.1f}%  {f}:{ln:<4d} {v[3]}\n")
        w.write("   top lines by shared-memory wavefronts (actual / ideal):\n")
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:14]:
            if v[1]:
                w.write(f"     {100 * v[1] / totw:5.1f}%  x{v[1] / max(1, v[2]):.2f}  {f}:{ln:<4d} {v[3]}\n")
        tots = sum(v[0] for v in samples.values()) or 1
        if samples:
            w.write("   top lines by warp-stall samples (where warps wait; top reasons):\n")
            for (f, ln), v in sorted(samples.items(), key=lambda kv: -kv[1][0])[:16]:
                w.write(f"     {100 * v[0] / tots:5.1f}%  {f}:{ln:<4d} {v[2][:70]}  [" + ", ".join(f"{a}={100 * b / tots:.1f}%" for a, b in v[1].most_common(3)) + "]\n")
    open(out, "w").write(w.getvalue())
    print(w.getvalue())


if __name__ == "__main__":
    main()
