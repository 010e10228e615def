"""ncu launch-list CSV (--metrics gpu__time_duration.sum --csv) -> profiles/<tag>_step_launches.csv (id,kernel,duration_us)
plus a per-library summary on stdout.   usage: python tools/launch_list.py gpurun_out/launches_step_X.csv profiles/r1_step_launches.csv
"""
import csv
import sys


def main(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    head = next(rd)
    kn, mv, mu = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    for r in rd:
        v = float(r[mv].replace(",", ""))
        us = v / 1e3 if r[mu] in ("ns", "nsecond") else v
        rows.append((len(rows), r[kn].replace(",", ";"), us))
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python bench.py --profile-step\n")
        f.write("# one eager chairs_uflow train step (B=8, 384x512, channels-last conv stacks) on B200, end of the round named in the file; "
                "per-launch times are cold-cache and serialised (compare shares)\n")
        f.write("id,kernel,duration_us\n")
        for i, k, us in rows:
            f.write("%d,%s,%.2f\n" % (i, k, us))
    tot = sum(r[2] for r in rows)
    groups = {"cudnn": 0.0, "arflow_b200": 0.0, "aten": 0.0}
    for _, k, us in rows:
        if "at::" in k[:40]:
            groups["aten"] += us
        elif "<unnamed>::" in k or "pad_weight_kernel" in k:
            groups["arflow_b200"] += us
        else:
            groups["cudnn"] += us
    print("%d launches, %.2f ms" % (len(rows), tot / 1e3))
    for g, us in groups.items():
        print("  %-12s %8.1f us  %5.1f %%" % (g, us, 100 * us / tot))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
