#!/usr/bin/env python
"""Per-source-line summary of an ncu report (no GPU needed): warp-stall samples and executed warp instructions of one
kernel, attributed to source lines through the line table of the matching object file.

    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep object_detectors_b200/csrc/rpn.o k_rpn_select_cluster [top]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[hdr_i - 1][:2] if hdr_i else "")
    hdr = rows[hdr_i]
    idx = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or r[0] == "Address" or not r[0].startswith("0x") and not re.match(r"^[0-9a-f]+$", r[0]):
            if r and r[0] == "Kernel Name":
                break
            continue
        data.append(r)
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
    cur, off2line = None, {}
    for l in sass[start + 1:]:
        if l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+", l)
        if m and cur:
            off2line[int(m.group(1), 16)] = cur
    base = min(int(r[idx["Address"]], 16) for r in data)
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in data:
        key = off2line.get(int(r[idx["Address"]], 16) - base, ("?", 0))
        a = agg[key]
        a[0] += int(r[idx["# Samples"]] or 0)
        a[1] += int(r[idx["Instructions Executed"]] or 0)
        for s_ in stalls:
            v = r[idx[s_]]
            if v and v != "0":
                a[2][s_[6:]] += int(v)
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print(f"samples {ts}  warp-instructions {ti}")
    srcdir = os.path.dirname(os.path.abspath(obj))
    cache = {}
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in cache:
            pth = os.path.join(srcdir, f)
            cache[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
        text = cache[f][ln - 1].strip()[:90] if 0 < ln <= len(cache[f]) else ""
        st = " ".join(f"{k}={c}" for k, c in v[2].most_common(2))
        print(f"{v[0]:6d} {100 * v[0] / ts:5.1f}% | instr {v[1]:9d} {100 * v[1] / ti:5.1f}% | {f}:{ln:<4d} {text}  [{st}]")


if __name__ == "__main__":
    main()
