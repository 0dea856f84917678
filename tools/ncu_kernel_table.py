"""One markdown row per profiled launch of an `ncu --set full` report: duration, DRAM bytes and %, tensor-pipe %, L1TEX data-pipe
shares (LSU shared wavefronts / tensor-core operand wavefronts), issue-slot utilisation, registers and the dominant warp stall.
usage: python tools/ncu_kernel_table.py report.ncu-rep"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]


def col(name):
    return hdr.index(name) if name in hdr else None


def val(r, name, default=""):
    i = col(name)
    if i is None or i >= len(r) or r[i] in ("", "no data"):
        return default
    return r[i]


stall_cols = [(i, re.sub(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", r"\1", h)) for i, h in enumerate(hdr)
              if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", h)]
print("| kernel | us | dram rd MB | dram wr MB | DRAM % | tensor % | LSU-smem wf % | TC-smem wf % | issue % | regs | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in data:
    name = re.sub(r"\(.*", "", val(r, "Kernel Name")).replace("void ", "").replace("tn::<unnamed>::", "").replace("<unnamed>::", "")

    def f(metric, scale=1.0, fmt="%.1f"):
        v = val(r, metric)
        try:
            return fmt % (float(v.replace(",", "")) * scale)
        except ValueError:
            return "-"
    stalls = []
    for i, n in stall_cols:
        try:
            stalls.append((float(r[i]), n))
        except (ValueError, IndexError):
            pass
    stalls = ", ".join("%s %.1f" % (n, v) for v, n in sorted(stalls, reverse=True)[:3] if n != "selected")
    unit_rd = rows[1][col("dram__bytes_read.sum")] if col("dram__bytes_read.sum") is not None else "Mbyte"
    sc = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6}.get(unit_rd, 1.0)
    unit_wr = rows[1][col("dram__bytes_write.sum")] if col("dram__bytes_write.sum") is not None else "Mbyte"
    scw = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6}.get(unit_wr, 1.0)
    tu = rows[1][col("gpu__time_duration.sum")]
    ts = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(tu, 1.0)
    print("| %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        name[:60], f("gpu__time_duration.sum", ts), f("dram__bytes_read.sum", sc), f("dram__bytes_write.sum", scw),
        f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        f("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        f("smsp__issue_active.avg.pct_of_peak_sustained_active"), f("launch__registers_per_thread", 1.0, "%.0f"), stalls))
