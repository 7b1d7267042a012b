"""ncu launch-list CSV (--metrics gpu__time_duration.sum --csv) -> markdown table of kernel shares.
    python tools/summarize_launches.py gpurun_out/launches.csv "title" > profiles/xxx.md"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*\)$", "", r["Kernel Name"]).replace("ssb::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        rows.append((name, us))
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}\n")
    print(f"Total kernel time {tot / 1000.0:.1f} ms over {len(rows)} launches (ncu serialises launches and runs them cold: read the SHARES).\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {us:.0f} | {100.0 * us / tot:.1f}% | {us / c:.1f} |")


if __name__ == "__main__":
    main()
