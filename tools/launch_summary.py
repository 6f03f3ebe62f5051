"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
launches, mean duration, share of the profiled time and DRAM bytes per launch for every (kernel, grid).
usage: python tools/launch_summary.py profiles/r01_launches_bench.csv"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr, rows = rows[0], rows[1:]
    ik, im, iv, iid, ig = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(r[iid], {"k": r[ik], "g": r[ig]})[r[im]] = float(r[iv].replace(",", ""))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        a = agg[(d["k"], d["g"])]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':72s} {'grid':18s} {'n':>4s} {'avg ms':>9s} {'share':>6s} {'rd GB':>8s} {'wr GB':>8s}")
    for (k, g), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} {g:18s} {a[0]:4d} {a[1] / a[0] / 1e6:9.3f} {100 * a[1] / tot:5.1f}% {a[2] / a[0] / 1e9:8.2f} {a[3] / a[0] / 1e9:8.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
