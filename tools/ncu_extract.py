"""Summarise .ncu-rep captures into a small CSV for profiles/ (run here: ncu reads reports without a GPU).
usage: python tools/ncu_extract.py out.csv rep1.ncu-rep [rep2 ...]"""
import csv, subprocess, sys
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_tmem.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__inst_executed_pipe_xu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_utc.sum"]
out = csv.writer(open(sys.argv[1], "w", newline=""))
first = True
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [h for h in WANT if h in hdr]
    if first:
        out.writerow(["report"] + cols)
        out.writerow(["(unit)"] + ["us" if c == "gpu__time_duration.sum" else ("byte" if "byte" in units[hdr.index(c)] else units[hdr.index(c)]) for c in cols])
        first = False
    SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        vals = []
        for c in cols:
            v, un = r[hdr.index(c)], units[hdr.index(c)]
            if c == "gpu__time_duration.sum":
                v = "%.3f" % (float(v) * SCALE[un])                      # always microseconds
            elif un in ("byte", "Kbyte", "Mbyte", "Gbyte") and v not in ("", "n/a"):
                v = "%.0f" % (float(v) * SCALE[un])                      # always bytes
            vals.append(v)
        out.writerow([rep.split("/")[-1]] + vals)
