"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, time, share."""
import collections, csv, sys

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot/1e6:.3f} ms total (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:64]:64s} n={v[0]:5d} total={v[1]/1e3:11.1f}us avg={v[1]/v[0]/1e3:9.1f}us share={100*v[1]/tot:5.1f}%")

if __name__ == "__main__":
    main(sys.argv[1])
