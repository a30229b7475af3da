"""Per-launch table from an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv`.

    ncu -i gpurun_out/conv_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_full.py /tmp/raw.csv [--json out.json]

Columns: duration, DRAM bytes read+written (the `traffic` of bench.py's roofline object), DRAM throughput % of
peak, tensor-pipe active % (sm__pipe_tensor_cycles_active), achieved occupancy, registers."""
import csv
import json
import re
import sys


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1], newline="") if not l.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}

    def find(pat):
        for n in hdr:
            if re.search(pat, n):
                return n
        return None

    names = dict(
        dur=find(r"^gpu__time_duration\.sum$"),
        rd=find(r"^dram__bytes_read\.sum$"), wr=find(r"^dram__bytes_write\.sum$"),
        dram_pct=find(r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$"),
        tensor=find(r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active$")
        or find(r"sm__pipe_tensor.*cycles_active.*pct"),
        inst_tensor=find(r"^sm__inst_executed_pipe_tensor.*pct"),
        regs=find(r"^launch__registers_per_thread$"), sm_clk=find(r"^sm__cycles_elapsed\.avg\.per_second$"),
        l2_hit=find(r"^lts__t_sector_hit_rate\.pct$"),
    )
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = []
    for r in data:
        def val(key):
            n = names.get(key)
            if n is None:
                return None
            v = num(r[col[n]])
            if v is None:
                return None
            return v * scale.get(units[col[n]], 1.0) if key in ("dur", "rd", "wr") else v
        name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]).replace("void ", "").replace("srgd::", "")
        out.append(dict(id=r[col["ID"]], kernel=name[:48], grid=r[col.get("Grid Size", 0)], us=val("dur"),
                        dram_bytes=(val("rd") or 0) + (val("wr") or 0), dram_pct=val("dram_pct"),
                        tensor_pct=val("tensor"), l2_hit=val("l2_hit"), regs=val("regs")))
    print(f"{'id':>4s} {'kernel':48s} {'us':>8s} {'dram MB':>9s} {'dram%':>6s} {'tensor%':>8s} {'L2hit%':>7s} {'regs':>5s}")
    for o in out:
        f = lambda v, w, p=1: (f"{v:{w}.{p}f}" if v is not None else " " * (w - 1) + "-")
        print(f"{o['id']:>4s} {o['kernel']:48s} {f(o['us'], 8)} {f(o['dram_bytes'] / 1e6, 9)} {f(o['dram_pct'], 6)} "
              f"{f(o['tensor_pct'], 8)} {f(o['l2_hit'], 7)} {f(o['regs'], 5, 0)}")
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
