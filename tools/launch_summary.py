"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.

    python tools/launch_summary.py gpurun_out/r02_launches_bench_timed_region.csv > profiles/r02_launches_bench_summary.csv
"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if r]
    hdr = rows[0]
    kname, mval = hdr.index("Kernel Name"), hdr.index("Metric Value")
    munit = hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= mval:
            continue
        v = float(r[mval].replace(",", ""))
        v = v / 1e3 if r[munit] in ("ns", "nsecond") else (v * 1e3 if r[munit] in ("ms", "msecond") else v)   # -> us
        name = re.sub(r"\(.*", "", r[kname]).strip()
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print("# per-kernel totals of the launch list (us); share of all profiled launches")
    print("kernel,launches,total_us,share")
    for k, v in tot.most_common():
        print(f"{k},{cnt[k]},{v:.1f},{v / total:.4f}")


if __name__ == "__main__":
    main()
