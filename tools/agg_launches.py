"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel time for ONE step
(the launches between two consecutive k_pack_weights_batch, the first kernel of l3_forward_backward, i.e. one step + Adam)."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"])
            for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]


def main(path, which=3, detail=False):
    rows = load(path)
    idx = [i for i, (n, t, g) in enumerate(rows) if "k_pack_weights_batch" in n]
    step = rows[idx[which]:idx[which + 1]]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, g in step:
        k = re.sub(r"\(.*", "", n).replace("void ", "").replace("l3::", "")
        agg[k][0] += 1
        agg[k][1] += t / 1e3
        if detail:
            print("%-44s %9.1f us  grid %s" % (k, t / 1e3, g))
    tot = sum(v[1] for v in agg.values())
    print("launches per step: %d   sum of kernel durations: %.1f us" % (len(step), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-46s n=%3d %9.1f us %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3, "--detail" in sys.argv)
