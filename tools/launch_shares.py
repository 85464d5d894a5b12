"""Per-kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised:
compare SHARES with the CUDA-event shares of bench.py, not absolutes)."""
import collections, csv, io, re, sys

def shares(fn, skip_until=None):
    lines = [l for l in open(fn) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void (mhdf::)?", "", r["Kernel Name"].split("(")[0])
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1000.0 if r["Metric Unit"].startswith("n") else (v * 1000.0 if r["Metric Unit"].startswith("m") else v)
        agg[name][0] += 1
        agg[name][1] += v
    return agg

if __name__ == "__main__":
    agg = shares(sys.argv[1])
    step = {k: v for k, v in agg.items() if k.startswith(("k_pass", "k_xfused", "k_spectral"))}
    tot = sum(v[1] for v in step.values())
    print(f"{'kernel':62s} {'launches':>8s} {'total us':>10s} {'share of step kernels':>22s}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        sh = f"{v[1] / tot:.3f}" if k in step else "-"
        print(f"{k[:62]:62s} {v[0]:8d} {v[1]:10.1f} {sh:>22s}")
