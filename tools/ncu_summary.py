"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics the roofline needs."""
import csv
import subprocess
import sys

KEYS = [("time_us", "gpu__time_duration.sum", 1e3),
        ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
        ("dram_read_MB", "dram__bytes_read.sum", 1),
        ("dram_write_MB", "dram__bytes_write.sum", 1),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("xbar2l1_TBps", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", 1),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1),
        ("grid", "launch__grid_size", 1)]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s  (units as reported by ncu: %s)" % (rep, ", ".join("%s[%s]" % (k, units[hdr.index(m)]) for k, m, _ in KEYS if m in hdr)))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("%-44s %s" % (d["Kernel Name"][:44], "  ".join("%s=%s" % (k, d.get(m, "?")[:10]) for k, m, _ in KEYS)))


if __name__ == "__main__":
    main(sys.argv[1])
