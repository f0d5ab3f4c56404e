"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
device time and share.  Usage: python tools/launch_summary.py gpurun_out/x_launches.csv [skip_first_n]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.DictReader(lines))[skip:]
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e3:.1f} us total device time (cold-cache, serialised)")
    print(f"# {'us':>10} {'n':>5} {'share':>6}  kernel")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e3:12.1f} {c:5d} {100 * t / tot:5.1f}%  {name[:110]}")


if __name__ == "__main__":
    main()
