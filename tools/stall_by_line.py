"""Attribute ncu warp-stall samples (source page, SASS level) to CUDA source lines.

    ncu -i X.ncu-rep --page source --csv --kernel-id :::N > src.csv
    cuobjdump -xelf all libaewn.so ; nvdisasm -g <file>.cubin > lines.txt
    python tools/stall_by_line.py src.csv lines.txt <mangled kernel name> [top]
"""
import collections
import csv
import os
import re
import sys


def main():
    src_csv, lines_txt, kname = sys.argv[1:4]
    top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    cur, off2line, inside = None, {}, False
    for ln in open(lines_txt):
        if ln.startswith(".text."):
            inside = ln.strip().rstrip(":") == ".text." + kname
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            off2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    base = int(data[0][ix["Address"]], 16)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(collections.Counter)
    for r in data:
        key = off2line.get(int(r[ix["Address"]], 16) - base)
        agg[key]["n"] += int(r[ix["# Samples"]])
        for h in stall_cols:
            agg[key][h[6:]] += int(r[ix[h]])
    cache = {}
    tot = sum(v["n"] for v in agg.values())
    print("total samples", tot)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top_n]:
        text = ""
        if k and os.path.exists(k[0]):
            if k[0] not in cache:
                cache[k[0]] = open(k[0]).read().split("\n")
            text = cache[k[0]][k[1] - 1].strip()[:90]
        top = {a: b for a, b in v.items() if a != "n" and b > 0.15 * v["n"]}
        print(f"{os.path.basename(k[0]) if k else None}:{k[1] if k else ''}", v["n"], f"{100 * v['n'] / tot:.1f}%", top, "|", text)


if __name__ == "__main__":
    main()
