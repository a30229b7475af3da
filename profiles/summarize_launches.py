"""Fold an `ncu --metrics gpu__time_duration.sum[,...] --csv` launch list into a per-kernel table.

    python profiles/summarize_launches.py gpurun_out/launches.csv [--per-launch]

ncu per-launch times are cold-cache and serialised: compare SHARES with the live CUDA-event numbers of
bench.py, not absolutes (B200_PROFILING.md)."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path, newline="") if not l.startswith("==")]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = launches.setdefault(int(r["ID"]), dict(name=r["Kernel Name"], grid=r["Grid Size"], block=r["Block Size"]))
        try:
            val = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r["Metric Name"]] = val * scale
    return launches


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("srgd::", "")
    return name[:60]


def main():
    path = sys.argv[1]
    per_launch = "--per-launch" in sys.argv
    L = load(path)
    total = sum(d.get("gpu__time_duration.sum", 0.0) for d in L.values())
    if per_launch:
        for i, d in L.items():
            t = d.get("gpu__time_duration.sum", 0.0)
            by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            print(f"{i:4d} {t:9.1f} us {100 * t / total:5.1f}%  dram {by / 1e6:8.1f} MB  grid {d['grid']:>16s} {short(d['name'])}")
        return
    agg = collections.OrderedDict()
    for d in L.values():
        a = agg.setdefault(short(d["name"]), dict(n=0, t=0.0, by=0.0))
        a["n"] += 1
        a["t"] += d.get("gpu__time_duration.sum", 0.0)
        a["by"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    if "--json" in sys.argv:
        import json
        out = sys.argv[sys.argv.index("--json") + 1]
        json.dump({"source": path, "note": "ncu launch list, one bench step (cold-cache, serialised): compare shares",
                   "total_ms": total / 1e3,
                   "kernels": {k: {"launches": a["n"], "ms": a["t"] / 1e3, "share": a["t"] / total,
                                   "dram_bytes": a["by"], "dram_bytes_per_launch": a["by"] / a["n"]}
                               for k, a in agg.items()}}, open(out, "w"), indent=1)
    print(f"{len(L)} launches, {total / 1e3:.3f} ms total (serialised, cold-cache)")
    print(f"{'kernel':60s} {'n':>4s} {'ms':>8s} {'share':>6s} {'dram GB':>8s} {'GB/s':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        gbs = a["by"] / (a["t"] * 1e-6) / 1e9 if a["t"] else 0.0
        print(f"{k:60s} {a['n']:4d} {a['t'] / 1e3:8.3f} {100 * a['t'] / total:5.1f}% {a['by'] / 1e9:8.3f} {gbs:7.0f}")


if __name__ == "__main__":
    main()
