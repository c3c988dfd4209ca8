"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_breakdown.py launches.csv > profiles/<name>.md"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.replace("b200::_GLOBAL__N__", "").replace("(anonymous namespace)::", "")
    return re.sub(r"^[0-9a-f_]+::", "", name)[:110]


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        a = agg.setdefault(short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.0f | %.1f | %.1f %% |" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / total))
    print("| **all** | %d | %.0f | %.1f | 100 %% |" % (n, total, total / max(n, 1)))


if __name__ == "__main__":
    main()
