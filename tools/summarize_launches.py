"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals/shares and per-(kernel,grid)."""
import collections
import csv
import re
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    shapes = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit == "ns" else v / 1e3 if unit.startswith("us") else v
        short = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        agg[short][0] += 1
        agg[short][1] += v
        shapes[(short[:44], row["Grid Size"])][0] += 1
        shapes[(short[:44], row["Grid Size"])][1] += v
    tot = sum(v[1] for v in agg.values())
    o = [f"# source: {path}", f"total {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches "
         "(ncu per-launch times are cold-cache and serialised: compare SHARES)", ""]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        o.append(f"{v[1]:9.2f} ms {100 * v[1] / tot:5.1f}% n={v[0]:4d}  {k}")
    o.append("")
    for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1])[:45]:
        o.append(f"{v[1]:9.2f} ms n={v[0]:3d} avg={v[1] / v[0]:.3f}  {k[0]} grid={k[1]}")
    txt = "\n".join(o)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
