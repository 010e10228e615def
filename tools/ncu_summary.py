"""Summarise an .ncu-rep (read here, no GPU needed): key metrics, stall reasons, hottest instructions.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--top 12] [--out profiles/x.txt]"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed.sum", "lts__t_bytes.sum.per_second", "l1tex__t_bytes.sum.per_second"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    out = open(sys.argv[sys.argv.index("--out") + 1], "w") if "--out" in sys.argv else None

    def emit(*a):
        line = " ".join(str(x) for x in a)
        print(line)
        if out:
            out.write(line + "\n")

    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    h, units, v = raw[0], raw[1], raw[2]
    emit("# ncu --set full --clock-control none summary of", rep)
    emit("kernel:", v[h.index("Kernel Name")][:120])
    for k in KEYS:
        if k in h:
            emit("%-70s %s %s" % (k, v[h.index(k)], units[h.index(k)]))
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    hdr = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    hh, data = src[hdr], []
    for r in src[hdr + 1:]:        # one launch only: stop at the next launch's header block
        if r and r[0] in ("Address", "Kernel Name"):
            break
        if len(r) == len(hh):
            data.append(r)
    si, ie = hh.index("# Samples"), hh.index("Instructions Executed")
    stalls = [i for i, x in enumerate(hh) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[si] or 0) for r in data) or 1
    emit("\n# warp-state samples: %d" % tot)
    agg = {hh[i]: sum(int(r[i] or 0) for r in data) for i in stalls}
    for k, val in sorted(agg.items(), key=lambda kv: -kv[1])[:9]:
        emit("  %-28s %5.1f%%" % (k, 100.0 * val / tot))
    mix = Counter()
    for r in data:
        toks = r[1].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        mix[op.split(".")[0]] += int(r[ie] or 0)
    t = sum(mix.values()) or 1
    emit("\n# executed instruction mix")
    emit("  " + "  ".join("%s %.1f%%" % (k, 100.0 * c / t) for k, c in mix.most_common(10)))
    emit("\n# hottest instructions (samples, top stall reasons)")
    for r in sorted(data, key=lambda r: -int(r[si] or 0))[:top]:
        st = sorted(((hh[i], int(r[i] or 0)) for i in stalls if int(r[i] or 0) > 0), key=lambda kv: -kv[1])[:2]
        emit("  %5s  %-58s %s" % (r[si], r[1].strip()[:58], st))
    if out:
        out.close()


if __name__ == "__main__":
    main()
