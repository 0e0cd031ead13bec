"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel microseconds per step."""
import collections
import csv
import sys


def main(path, steps):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    dur, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0][-70:]
        dur[k] = dur.get(k, 0.0) + v
        cnt[k] += 1
    tot = sum(dur.values())
    for k, v in sorted(dur.items(), key=lambda kv: -kv[1]):
        print(f"{v / steps / 1000:9.1f} us/step  {100 * v / tot:5.1f} %  x{cnt[k] / steps:.0f}  {k}")
    print(f"{tot / steps / 1000:9.1f} us/step total, {sum(cnt.values()) / steps:.0f} launches/step")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
