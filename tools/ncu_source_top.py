"""Top stall lines of an `ncu --set full --import-source on` report: python tools/ncu_source_top.py rep.ncu-rep [n]"""
import csv
import io
import subprocess
import sys


def main(path, n=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    key = "Warp Stall Sampling (All Samples)"
    tot = sum(int(r[key] or 0) for r in rows)
    print("total samples", tot, "instructions", len(rows))
    ranked = sorted(enumerate(rows), key=lambda ir: -int(ir[1][key] or 0))[:n]
    for i, r in ranked:
        print("%5d %6.2f%%  exec=%-8s %s" % (i, 100.0 * int(r[key] or 0) / max(tot, 1), r["Instructions Executed"], r["Source"].strip()[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
