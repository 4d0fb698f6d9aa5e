"""`ncu -i report.ncu-rep --page raw --csv` (optionally .gz) -> the two summary tables kept under profiles/:
   <out>_all_kernels_ncu.csv  one row per profiled launch: ms, DRAM MB read / written, SM throughput, lanes per instruction, pipes, hit rates,
                              registers, DRAM GB/s
   <out>_trace8_ncu.csv       the k_trace8 launches only, in the column layout bench.py reads (roofline.traffic, issue_active, lanes_per_inst)
usage: ncu_summarize.py raw.csv[.gz] out_prefix"""
import csv, gzip, sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
TIME = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}


def num(x):
    try: return float(str(x).replace(",", ""))
    except ValueError: return float("nan")


def kernel_name(full):
    """'void k_trace8<(bool)0, (bool)0>(Frame, int, const unsigned int *)' -> 'k_trace8<0, 0>'"""
    t = full.replace("void ", "").replace("(bool)", "").replace("(int)", "").strip()
    depth = 0
    for i, ch in enumerate(t):
        if ch == "<": depth += 1
        elif ch == ">": depth -= 1
        elif ch == "(" and depth == 0: return t[:i].strip()
    return t


def main(path, out):
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = [r for r in csv.reader(f) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    ix = {n: i for i, n in enumerate(names)}
    kn = ix["Kernel Name"]
    reg = ix.get("launch__registers_per_thread")
    table = []
    for li, r in enumerate(data):
        if len(r) < len(names): continue
        vals = []
        for c in COLS:
            if c not in ix: vals.append(float("nan")); continue
            v, u = num(r[ix[c]]), units[ix[c]]
            if c == "gpu__time_duration.sum": v *= TIME.get(u, 1.0)
            elif c.startswith("dram__bytes"): v *= BYTES.get(u, 1.0)
            vals.append(v)
        name = kernel_name(r[kn])
        regs = int(num(r[reg])) if reg is not None else 0
        gbs = (vals[1] + vals[2]) / vals[0] if vals[0] > 0 else 0.0        # MB / ms = GB/s
        table.append((li, name, vals, regs, gbs))
    with open(out + "_all_kernels_ncu.csv", "w", newline="") as g:
        w = csv.writer(g); w.writerow(["launch", "kernel"] + COLS + ["registers", "dram_gbs"])
        for li, name, vals, regs, gbs in table:
            w.writerow([li, name] + [round(v, 4) for v in vals] + [regs, round(gbs, 1)])
    with open(out + "_trace8_ncu.csv", "w", newline="") as g:
        w = csv.writer(g); w.writerow(["launch", "kernel"] + COLS)
        k = 0
        for li, name, vals, regs, gbs in table:
            if not name.startswith("k_trace8"): continue
            w.writerow([k, name] + [round(v, 4) for v in vals]); k += 1
            if k == 8: break          # the wave frame's launches (the SVGF frame that follows traces one pass)
    print(f"{len(table)} launches summarised -> {out}_all_kernels_ncu.csv, {out}_trace8_ncu.csv")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
