"""Per-source-line and per-file share of warp-stall samples and executed instructions of one profiled kernel.
Run here (no GPU): python profiles/hotlines.py gpurun_out/x.ncu-rep [top N]
Needs a report captured with --import-source on from a library built with -lineinfo."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr = None, None
    by_line = defaultdict(lambda: [0, 0, ""])  # samples, instructions, text
    by_file = defaultdict(lambda: [0, 0])
    stall_by_file = defaultdict(lambda: defaultdict(int))
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name" or hdr is None or len(r) < len(hdr):
            continue
        if r[hdr.index("Address")] != "-":  # keep the CUDA source rows only (SASS rows repeat their numbers)
            continue
        try:
            line = int(r[0])
            samples = int(r[hdr.index("# Samples")] or 0)
            inst = int(r[hdr.index("Instructions Executed")] or 0)
        except ValueError:
            continue
        key = (cur_file, line)
        by_line[key][0] += samples
        by_line[key][1] += inst
        by_line[key][2] = r[1].strip()
        by_file[cur_file][0] += samples
        by_file[cur_file][1] += inst
        for name in ("stall_barrier", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_branch_resolving", "stall_long_sb",
                     "stall_math", "stall_not_selected", "stall_selected", "stall_mio", "stall_dispatch"):
            if name in hdr:
                stall_by_file[cur_file][name] += int(r[hdr.index(name)] or 0)
    ts = sum(v[0] for v in by_file.values()) or 1
    ti = sum(v[1] for v in by_file.values()) or 1
    print(f"total samples {ts}, warp instructions {ti}")
    print("\nby file: samples%  instr%   dominant stalls")
    for f, (s, i) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        st = sorted(stall_by_file[f].items(), key=lambda kv: -kv[1])[:4]
        print(f"  {f:16s} {100 * s / ts:6.1f} {100 * i / ti:7.1f}   " + ", ".join(f"{k[6:]} {100 * v / max(1, s):.0f}%" for k, v in st))
    print(f"\ntop {top} lines by samples: file:line samples% instr% source")
    for (f, ln), (s, i, text) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {f}:{ln:<5d} {100 * s / ts:5.1f} {100 * i / ti:5.1f}  {text[:110]}")


if __name__ == "__main__":
    main()
