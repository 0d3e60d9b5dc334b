"""Summarise an ncu launch list (gpu__time_duration.sum --csv) per kernel.   python tools/launch_summary.py file.csv [title]"""
import collections
import csv
import re
import sys


def summarize(path, title, out=sys.stdout):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    rows = [(re.sub(r"\(.*", "", row[ki]), float(row[vi].replace(",", "")) / 1e6) for row in r if len(row) > vi]
    tot, cnt, mx = collections.Counter(), collections.Counter(), collections.Counter()
    for k, ms in rows:
        tot[k] += ms
        cnt[k] += 1
        mx[k] = max(mx[k], ms)
    S = sum(tot.values())
    out.write(f"{title}\n{len(rows)} launches, {S:.2f} ms of kernel time (ncu gpu__time_duration.sum, --clock-control none; cold-cache, "
              f"serialised: compare SHARES, not absolutes)\n\n")
    out.write(f"{'kernel':42s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'max ms':>9s}\n")
    for k, v in tot.most_common():
        out.write(f"{k[:42]:42s} {cnt[k]:8d} {v:10.2f} {100 * v / S:6.1f}% {mx[k]:9.3f}\n")
    out.write("\n")


if __name__ == "__main__":
    summarize(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
