#!/usr/bin/env python3
"""Static evidence for the final build (no GPU needed): per kernel the ptxas line (registers, spills, barriers, static
shared memory) and a histogram of the SASS it was compiled to (cuobjdump -sass of wordpiece_b200/lib/obj/wp_encode.o):
instruction count, the widest global loads / stores, shared-memory, atomic, vote / shuffle / barrier and XU-pipe
(POPC, FLO, BREV) counts.     python tools/sass_summary.py > profiles/<tag>_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "wordpiece_b200", "lib", "obj", "wp_encode.o")
LOG = os.path.join(ROOT, "wordpiece_b200", "lib", "obj", "wp_encode.ptxas.log")


def demangle(name: str) -> str:
    try:
        return subprocess.check_output(["c++filt", name], text=True).strip()
    except Exception:  # noqa: BLE001
        return name


def ptxas_lines():
    out, cur = {}, None
    for line in open(LOG):
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        if cur and ("Used " in line or "spill" in line):
            out[cur].append(line.replace("ptxas info    :", "").strip())
    return out


def sass_histograms():
    txt = subprocess.check_output(["cuobjdump", "-sass", OBJ], text=True)
    hist, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            hist[cur][m.group(1)] += 1
    return hist


def main():
    px, hs = ptxas_lines(), sass_histograms()
    def width(prefix):
        def pats():
            return [prefix + r"(\.\w+)*\.256", prefix + r"(\.\w+)*\.128", prefix + r"(\.\w+)*\.64(\.|$)",
                    prefix + r"(?!.*\.(256|128|64|U8|S8|U16|S16)(\.|$))", prefix + r"(\.\w+)*\.[US](8|16)(\.|$)"]
        return pats()

    groups = [("global loads 256 / 128 / 64 / 32 / 8-16 bit", width(r"^LDG")),
              ("global stores 256 / 128 / 64 / 32 / 8-16 bit", width(r"^STG")),
              ("shared loads / stores", [r"^LDS", r"^STS"]),
              ("atomics + reductions global / shared", [r"^(ATOMG|REDG|RED|ATOM)(\.|$)", r"^ATOMS"]),
              ("barriers / votes / shuffles / warp reductions", [r"^BAR", r"^VOTE", r"^SHFL", r"^REDUX"]),
              ("XU pipe: POPC / FLO / BREV", [r"^POPC", r"^FLO", r"^BREV"]),
              ("local memory LDL / STL", [r"^LDL", r"^STL"])]
    print("Static summary of wordpiece_b200/lib/obj/wp_encode.o (nvcc -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a)\n")
    for fn in sorted(hs, key=lambda f: -sum(hs[f].values())):
        h = hs[fn]
        print(f"== {demangle(fn)}")
        for line in px.get(fn, []):
            print(f"   ptxas: {line}")
        print(f"   SASS instructions: {sum(h.values())}")
        for title, pats in groups:
            counts = [sum(c for op, c in h.items() if re.search(p, op)) for p in pats]
            print(f"   {title}: {' / '.join(map(str, counts))}")
        top = ", ".join(f"{op} {c}" for op, c in h.most_common(12))
        print(f"   most frequent: {top}\n")


if __name__ == "__main__":
    sys.exit(main())
