"""Summarise gpurun_out ncu artefacts into small text files under profiles/ (tracked).
usage: python tools/ncu_summary.py <launches.csv> <report.ncu-rep> <out-prefix>"""
import collections, csv, subprocess, sys, re

def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try: v = float(d["Metric Value"].replace(",", ""))
            except ValueError: continue
            unit = d["Metric Unit"]
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            agg[d["Kernel Name"]][0] += 1; agg[d["Kernel Name"]][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n# source: {path}\n")
        f.write(f"{'ms total':>12} {'launches':>8} {'share':>8}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:12.3f} {v[0]:8d} {100*v[1]/tot:7.2f}%  {k}\n")

KEEP = re.compile(r"^(gpu__time_duration.sum|launch__(registers_per_thread|occupancy_limit_.*|grid_size|block_size|waves_per_multiprocessor)|"
                  r"sm__warps_active.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum|"
                  r"sm__inst_executed_pipe_(xu|fma|alu|lsu|fp64|uniform|fmaheavy|fmalite).avg.pct_of_peak_sustained_active|sm__pipe_(fma|fmaheavy|fmalite|alu|fp64)_cycles_active.avg.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__throughput.avg.pct_of_peak_sustained_elapsed|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|dram__bytes_(read|write).sum|"
                  r"smsp__cycles_elapsed.avg.per_second|sm__cycles_elapsed.avg|smsp__warps_eligible.avg.per_cycle_active|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|"
                  r"sm__sass_thread_inst_executed_op_(fadd|fmul|ffma|fp32)_pred_on.sum|smsp__sass_thread_inst_executed_op_.*_pred_on.sum)$")

def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source: {rep}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n== {d.get('Kernel Name')}  grid={d.get('Grid Size')} block={d.get('Block Size')}\n")
            for h in hdr:
                if KEEP.match(h) and d[h] not in ("", "n/a"):
                    f.write(f"{h} = {d[h]} {units[hdr.index(h)]}\n")

if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[3] + "_launches.txt")
    full(sys.argv[2], sys.argv[3] + "_full.txt")
