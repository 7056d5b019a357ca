#!/usr/bin/env python
"""Attribute an ncu --set full report's warp-stall samples of one kernel to CUDA source lines.
ncu's CSV source page is SASS-only; nvdisasm -g supplies the SASS -> file:line map (same order).

    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-mangled-substring> [top] [ncu-kernel-regex]
(the last argument selects one kernel of a multi-kernel report, e.g. k_rows_finish)
"""
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def main(rep, lib, kern, top=40, kregex=None):
    sel = ["-k", "regex:" + kregex, "-c", "1"] if kregex else []
    raw = subprocess.run(["ncu", "-i", rep, *sel, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=tmp, capture_output=True)
    cubin = next(Path(tmp).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
    # instructions of the kernel section and of every function placed after it in ncu's listing
    lines = []  # (file, line) per instruction, in order
    cur = ("?", 0)
    insec = False
    for ln in dis:
        if ln.startswith("//--------------------- .text."):
            insec = kern in ln
            continue
        if not insec:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    ismp, iex, ith = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    n = min(len(lines), len(data))
    print(f"sass rows {len(data)}  disasm instrs {len(lines)}")
    agg = {}
    for k in range(n):
        r = data[k]
        a = agg.setdefault(lines[k], [0, 0, 0.0, {}])
        s = int(r[ismp] or 0)
        e = int(r[iex] or 0)
        a[0] += s
        a[1] += e
        a[2] += e * float(r[ith] or 0)
        for i in stall:
            v = int(r[i] or 0)
            if v:
                a[3][hdr[i][6:]] = a[3].get(hdr[i][6:], 0) + v
    tot = sum(a[0] for a in agg.values()) or 1
    tote = sum(a[1] for a in agg.values()) or 1
    src = {}
    print(f"samples {tot} warp-instructions {tote}")
    for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        if f not in src:
            cand = list(Path(lib).resolve().parent.rglob(f))
            src[f] = cand[0].read_text().splitlines() if cand else []
        text = src[f][l - 1].strip()[:80] if 0 < l <= len(src[f]) else ""
        st = sorted(a[3].items(), key=lambda x: -x[1])[:2]
        thr = a[2] / a[1] if a[1] else 0
        print(f"{f}:{l:<4} smp {a[0]/tot:5.1%} inst {a[1]/tote:5.1%} thr {thr:4.1f} {st} | {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40, sys.argv[5] if len(sys.argv) > 5 else None)
