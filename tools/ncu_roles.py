"""Per-role view of a warp-specialised kernel's ncu capture: lists the SASS landmarks (mbarrier waits / arrives, tcgen05
MMAs, TMEM loads, bulk copies) with their stall samples and execution counts, and sums the samples between cut points.

    python tools/ncu_roles.py capture.ncu-rep [cut,cut,...]     # cuts = SASS indices separating the roles
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    cuts = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else []
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def num(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0
    keys = ("UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "TRYWAIT", "ARRIVE", "UTMALDG", "UTMASTG", "EXIT", "BAR.SYNC")
    for i, r in enumerate(data):
        s = r[ix["Source"]]
        if any(k in s for k in keys):
            win = sum(num(data[j], "# Samples") for j in range(i, min(i + 4, len(data)))) if "TRYWAIT" in s else num(r, "# Samples")
            print(f"{i:5d} samples {int(win):6d} exec {int(num(r, 'Instructions Executed')):9d}  {' '.join(s.split())[:90]}")
    total = sum(num(r, "# Samples") for r in data)
    print("total samples", int(total))
    edges = [0] + cuts + [len(data)]
    for a, b in zip(edges[:-1], edges[1:]):
        print(f"[{a:5d},{b:5d})  samples {int(sum(num(data[i], '# Samples') for i in range(a, b))):6d}  "
              f"warp-instructions {int(sum(num(data[i], 'Instructions Executed') for i in range(a, b))):10d}")


if __name__ == "__main__":
    main()
