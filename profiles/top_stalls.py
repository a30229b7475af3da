"""Top stalled SASS instructions per kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source sass > /tmp/src.csv
    python profiles/top_stalls.py /tmp/src.csv [N]
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], hdr=None, rows=[])
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    for k in kernels:
        h = {n: i for i, n in enumerate(k["hdr"])}
        tot = sum(int(r[h["# Samples"]]) for r in k["rows"]) or 1
        ex = sum(int(r[h["Instructions Executed"]]) for r in k["rows"])
        print(f"===== {k['name'][:70]}: {tot} samples, {ex} warp-instructions, {len(k['rows'])} SASS lines")
        stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
        for i, r in sorted(enumerate(k["rows"]), key=lambda ir: -int(ir[1][h["# Samples"]]))[:top_n]:
            n = int(r[h["# Samples"]])
            why = sorted(((int(r[h[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
            why = " ".join(f"{c}:{v}" for v, c in why if v)
            print(f"{i:5d} {n:6d} {100 * n / tot:5.1f}% ex={r[h['Instructions Executed']]:>9s}  {r[h['Source']].strip()[:70]:70s} {why}")
        ops = collections.Counter()
        for r in k["rows"]:
            toks = [t for t in r[h["Source"]].strip().split() if not t.startswith("@")]
            if toks:
                ops[toks[0].split(".")[0]] += int(r[h["Instructions Executed"]])
        print("   mix:", ", ".join(f"{o} {100 * n / ex:.1f}%" for o, n in ops.most_common(14)))


if __name__ == "__main__":
    main()
