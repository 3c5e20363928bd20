"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals/shares."""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        out.append((row["Kernel Name"], row["Grid Size"], v))
    return out


def main():
    rows = load(sys.argv[1])
    detail = len(sys.argv) > 2
    tot, cnt = {}, collections.Counter()
    for name, grid, v in rows:
        key = name.split("(")[0]
        tot[key] = tot.get(key, 0) + v
        cnt[key] += 1
    T = sum(tot.values())
    print(f"total {T/1000:.3f} ms over {len(rows)} launches")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:28]:
        print(f"| `{k[:100]}` | {cnt[k]} | {v:.1f} | {100*v/T:.1f}% |")
    if detail:
        for name, grid, v in rows:
            if sys.argv[2] in name:
                print(name.split("(")[0][-30:], grid, f"{v:.1f}")


if __name__ == "__main__":
    main()
