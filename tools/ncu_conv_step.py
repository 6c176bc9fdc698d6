"""One training step's convolution launches from an `ncu --set full` report -> the per-launch table committed under
profiles/ and the per-class DRAM traffic / tensor-pipe summary bench.py quotes (JSON).

    python tools/ncu_conv_step.py gpurun_out/x.ncu-rep profiles/r1_ncu_full_conv_step.txt profiles/r1_ncu_conv_classes.json
"""
import csv
import json
import re
import subprocess
import sys


def klass(name):
    if "wgrad" in name or "bias_grad" in name:
        return "conv_wgrad"
    if "first_conv" in name:
        return "conv_fwd"
    m = re.search(r"k_conv3x3_tc3<\d+, \d+, (\d+)>", name)
    if m:
        # EPI_NONE (0) = no statistics, EPI_BWD (5) = BN-backward statistics of the layer below: data gradients
        return "conv_dgrad" if m.group(1) in ("0", "5") else "conv_fwd"
    return None


def main(rep, txt, js):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    units = rows[1]

    def val(r, key, want_unit):
        v = float(r[ix[key]].replace(",", ""))
        u = units[ix[key]]
        scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "us": 1.0, "ms": 1e3, "ns": 1e-3, "%": 1.0, "": 1.0}
        return v * scale.get(u, 1.0)

    lines = ["# ncu --set full --clock-control none, all convolution launches of ONE training step (B=64, bf16), per launch",
             "# kernel | time_us | tensor_pipe_active_% | dram_read_MB | dram_write_MB | grid"]
    cls = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("l3::", "")
        name = name.replace("(int)", "").replace("(bool)", "")
        t = val(r, "gpu__time_duration.sum", "us")
        tp = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "%") if \
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in ix else float("nan")
        rd, wr = val(r, "dram__bytes_read.sum", "Mbyte"), val(r, "dram__bytes_write.sum", "Mbyte")
        lines.append("%-34s %7.1f %7.1f %10.1f %10.1f  %s" % (name, t, tp, rd, wr, r[ix["launch__grid_size"]]))
        k = klass(name)
        if k:
            c = cls.setdefault(k, {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0, "tensor_time_us": 0.0})
            c["launches"] += 1
            c["time_us"] += t
            c["dram_bytes"] += (rd + wr) * 1e6
            c["tensor_time_us"] += t * tp / 100.0
    for c in cls.values():
        c["tensor_pipe_active_pct_time_weighted"] = 100.0 * c["tensor_time_us"] / c["time_us"]
        c["dram_bytes_per_launch"] = c["dram_bytes"] / c["launches"]
    open(txt, "w").write("\n".join(lines) + "\n")
    json.dump(cls, open(js, "w"), indent=1)
    print(json.dumps(cls, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
