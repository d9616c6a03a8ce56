"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (profiles/*_launch_shares.csv).

    python tools/launch_shares.py gpurun_out/r2b_launches_raw.csv > profiles/r2b_launch_shares.csv
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v
        name = re.sub(r"\(.*", "", r[ik])[:100]
        tot[name] += ms
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {sum(cnt.values())} launches, {total:.3f} ms (per-launch times under ncu are cold-cache and serialised: compare SHARES)")
    print("ms,share_pct,launches,kernel")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{tot[k]:.3f},{100 * tot[k] / total:.1f},{cnt[k]},{k}")


if __name__ == "__main__":
    main()
