"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: count, total, average, share."""
import collections, csv, re, sys

for path in sys.argv[1:]:
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, seq = collections.defaultdict(lambda: [0, 0.0]), []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:56]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
        agg[k][0] += 1
        agg[k][1] += v
        seq.append((k, v, row["Grid Size"], row["Block Size"]))
    tot = sum(v[1] for v in agg.values())
    print(f"{path}: {len(seq)} launches, {tot / 1e3:.2f} ms of device time")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
        print(f"  {k:56s} n={v[0]:5d} total={v[1] / 1e3:9.2f} ms  avg={v[1] / v[0]:9.1f} us  {100 * v[1] / tot:5.1f}%")
