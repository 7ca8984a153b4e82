#!/usr/bin/env python3
"""Per-source-line instruction and stall table of one kernel from an `ncu --set full --import-source on` capture.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top N]
Aggregates the cuda,sass source page by source line: warp instructions executed, share of the kernel,
average active threads, stall samples, shared-memory excess wavefronts.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern, "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
    rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    i_line, i_src = 0, 1
    i_inst = ix["Instructions Executed"]
    i_tinst = ix["Thread Instructions Executed"]
    i_samp = ix["# Samples"]
    i_exc = ix.get("L1 Wavefronts Shared Excessive")
    i_wav = ix.get("L1 Wavefronts Shared")
    agg = {}
    total = 0
    tsamp = 0
    for r in rows[1:]:
        if len(r) <= i_inst:
            continue
        try:
            ln = int(r[i_line])
            inst = int(r[i_inst] or 0)
        except ValueError:
            continue
        a = agg.setdefault(ln, {"src": r[i_src], "inst": 0, "tinst": 0, "samp": 0, "exc": 0, "wav": 0, "sass": 0})
        a["inst"] += inst
        a["tinst"] += int(r[i_tinst] or 0)
        a["samp"] += int(r[i_samp] or 0)
        a["sass"] += 1
        if i_exc is not None:
            a["exc"] += int(r[i_exc] or 0)
            a["wav"] += int(r[i_wav] or 0)
        total += inst
        tsamp += int(r[i_samp] or 0)
    print(f"kernel {kern}: {total} warp instructions, {tsamp} stall samples, {len(agg)} source lines")
    print(f"{'line':>5} {'inst%':>6} {'cum%':>6} {'thr/inst':>8} {'samp%':>6} {'sass':>5} {'smem exc/wav':>14}  source")
    cum = 0
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
        cum += a["inst"]
        print(f"{ln:5d} {100 * a['inst'] / total:6.2f} {100 * cum / total:6.1f} {a['tinst'] / max(a['inst'], 1):8.1f} "
              f"{100 * a['samp'] / max(tsamp, 1):6.2f} {a['sass']:5d} {a['exc']:>7}/{a['wav']:<7} {a['src'].strip()[:90]}")


if __name__ == "__main__":
    main()
