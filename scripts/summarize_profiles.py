"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/.

  launches <csv> <out.md>   per-kernel share of ONE train step (between two advance_step_kernel launches) from an
                            `ncu --metrics gpu__time_duration.sum --csv` launch list
  rep <file.ncu-rep> <out.csv> [kernel-substring]   key metrics of the first matching kernel of a `--set full` capture
"""
import csv, re, subprocess, sys


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    r = list(csv.reader(lines)); hdr = r[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = [(x[ki], float(x[vi].replace(",", ""))) for x in r[1:]]
    starts = [i for i, d in enumerate(data) if "advance_step" in d[0]]
    step = data[starts[-2]:starts[-1]]
    tot = sum(d[1] for d in step)
    agg = {}
    for k, v in step:
        k = re.sub(r"\(.*", "", k).replace("void ", "")
        k = re.sub(r"CUB_\w+::", "cub::", k)[:90]
        a = agg.setdefault(k, [0.0, 0]); a[0] += v; a[1] += 1
    with open(dst, "w") as f:
        f.write(f"source: {src}\none train step = {len(step)} launches, {tot / 1e3:.1f} us summed kernel time "
                "(ncu: serialised, cold caches - use the SHARES, not the absolutes)\n\n| kernel | launches | us | share |\n|---|---|---|---|\n")
        for k, (v, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
            f.write(f"| `{k}` | {n} | {v / 1e3:.1f} | {100 * v / tot:.1f}% |\n")


KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "launch__waves_per_multiprocessor"]


def rep(src, dst, match=""):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
    row = next(r for r in rows[2:] if match in r[hdr.index("Kernel Name")])
    with open(dst, "w") as f:
        w = csv.writer(f); w.writerow(["metric", "value", "unit"])
        for k in KEYS:
            if k in hdr:
                w.writerow([k, row[hdr.index(k)], units[hdr.index(k)]])


def rep_table(src, dst, peak_gbs="6553.6"):
    """One line per captured launch of a `--set full` capture: duration, DRAM bytes, achieved DRAM GB/s and its fraction
    of the measured copy peak, occupancy - the HBM-bound row kernels of a train step."""
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
    col = lambda r, k: r[hdr.index(k)] if k in hdr else ""
    num = lambda x: float(x.replace(",", "")) if x else 0.0
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    with open(dst, "w") as f:
        f.write(f"source: {src} (ncu --set full --clock-control none; cold-ish caches, each launch replayed)\n"
                f"peak = {peak_gbs} GB/s (MEASURED_PEAKS.json copy bandwidth)\n\n"
                "| kernel | grid | us | DRAM read MB | DRAM write MB | DRAM GB/s | of peak | dram pct (ncu) | warps active % | regs |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", col(r, "Kernel Name")).replace("void ", "")[:60]
            t = num(col(r, "gpu__time_duration.sum")) * tscale.get(units[hdr.index("gpu__time_duration.sum")], 1e-9)
            rd = num(col(r, "dram__bytes_read.sum")) * scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
            wr = num(col(r, "dram__bytes_write.sum")) * scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
            gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
            f.write(f"| `{name}` | {col(r, 'Grid Size')} | {t * 1e6:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {gbs / float(peak_gbs):.2f} | "
                    f"{col(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed')} | {col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | "
                    f"{col(r, 'launch__registers_per_thread')} |\n")


if __name__ == "__main__":
    {"launches": launches, "rep": rep, "rep_table": rep_table}[sys.argv[1]](*sys.argv[2:])
