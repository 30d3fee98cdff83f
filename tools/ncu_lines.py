#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel in an .ncu-rep (captured with --import-source on,
compiled with -lineinfo):   python tools/ncu_lines.py REPORT KERNEL_REGEX [TOP]
Development tool: reads `ncu --page source --print-source cuda,sass --csv`."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    fname, hdr, agg, seen = None, None, {}, set()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            if fname in seen:                      # a second launch of the kernel: the first one is enough
                break
            seen.add(fname)
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0]:                                   # a source line row: start of a group
            cur = (fname, int(r[0]), r[1].strip()[:90])
            agg.setdefault(cur, [0, 0])
        if r[hdr["Address"]]:
            try:
                agg[cur][0] += int(r[hdr["Instructions Executed"]] or 0)
                agg[cur][1] += int(r[hdr["# Samples"]] or 0)
            except ValueError:
                pass
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {ti}, samples {ts}")
    for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*i/ti:5.1f}% inst {100*s/ts:5.1f}% smp  {f}:{ln:<4d} {src}")


if __name__ == "__main__":
    main()
