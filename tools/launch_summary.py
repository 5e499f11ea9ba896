"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections, csv, re, sys

def main(fn, top=30):
    with open(fn) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    vals = [float(r["Metric Value"]) for r in rows]
    tot = sum(vals)
    agg = collections.OrderedDict()
    for n, v in zip(names, vals):
        k = re.sub(r"\(.*", "", n)
        k = re.sub(r"^void ", "", k)
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    print(f"{len(rows)} launches, total {tot / 1e6:.3f} ms (cold-cache, serialised: compare shares, not absolutes)")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{t / 1e3:10.1f} us  {100 * t / tot:5.1f}%  x{c:3d}  {k[:90]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
