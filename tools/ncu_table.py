"""Summarise `ncu --set full` reports (one kernel each) into one CSV row per report."""
import csv, io, subprocess, sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]

out = csv.writer(sys.stdout)
first = True
first_units = None
for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT]
    if first:
        out.writerow(WANT)
        out.writerow([units[i] for i in idx])
        first = False
    # ncu picks a unit per report (us / ms, Kbyte / Mbyte ...): convert to the units of the first report's header row
    scale = {("us", "ms"): 1e-3, ("ms", "us"): 1e3, ("ns", "ms"): 1e-6, ("ns", "us"): 1e-3, ("s", "ms"): 1e3,
             ("Kbyte", "Mbyte"): 1e-3, ("Mbyte", "Kbyte"): 1e3, ("byte", "Mbyte"): 1e-6, ("byte", "Kbyte"): 1e-3,
             ("Gbyte", "Mbyte"): 1e3, ("Kbyte", "byte"): 1e3, ("Mbyte", "byte"): 1e6}
    if first_units is None:
        first_units = [units[i] for i in idx]
    for r in rows[2:]:
        vals = []
        for k, i in enumerate(idx):
            v, u, u0 = r[i], units[i], first_units[k]
            key = (u.split('/')[0], u0.split('/')[0])
            if u != u0 and key in scale and u.split('/')[1:] == u0.split('/')[1:]:
                try:
                    v = "%.6f" % (float(v.replace(",", "")) * scale[key])
                except ValueError:
                    pass
            vals.append(v)
        out.writerow(vals)
