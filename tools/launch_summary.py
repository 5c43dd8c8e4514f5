"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total and mean ms."""
import collections, csv, sys

def summarise(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:48]
        a = agg.setdefault(k, [0, 0.0, r[gi]])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t, _ in agg.values())
    print("kernel,launches,total_ms,mean_ms,share,first_grid")
    for k, (n, t, g) in agg.items():
        print(f"{k},{n},{t / 1e6:.3f},{t / n / 1e6:.4f},{t / tot:.3f},{g}")

if __name__ == "__main__":
    summarise(sys.argv[1])
