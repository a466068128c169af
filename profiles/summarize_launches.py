"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share of ONE step.
usage: python profiles/summarize_launches.py launches.csv [step_index_from_end=1]"""
import collections
import csv
import re
import sys


def main(path, from_end=1):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    data = rows[hi + 1:]
    names = [r[ki] for r in data]
    # 5 coordinate maps are built per step: steps are delimited by every 5th insert_kernel launch
    ins = [i for i, n in enumerate(names) if "insert_kernel" in n]
    starts = [ins[i] for i in range(0, len(ins), 5)]
    steps = []
    for a, b in zip(starts, starts[1:] + [len(names)]):
        if any("conv" in n or "wgrad" in n for n in names[a:b]):   # training steps, not the map-build timing loop
            steps.append((a - 1, b - 1))
    a, b = steps[-from_end]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data[a:b]:
        n = re.sub(r"<.*", "", r[ki]).replace("void ", "")
        n = re.sub(r"\(.*", "", n)
        agg[n][0] += 1
        agg[n][1] += float(r[vi]) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f"step launches {b - a}, summed kernel time {tot:.2f} ms (cold-cache, serialised under ncu: compare SHARES)")
    print(f"{'ms':>10} {'share':>6} {'launches':>8}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
        print(f"{t:10.3f} {100 * t / tot:5.1f}% {c:8d}  {n[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
