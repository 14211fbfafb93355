"""Summarises an .ncu-rep (run here, no GPU needed): headline metrics, stall mix and the
hottest SASS instructions of the first captured kernel.  usage: ncu_summary.py REPORT [TOPN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
hdr, units, rows = raw[0], raw[1], raw[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for r in rows:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-62s %s %s" % (w, r[i], units[i]))
    print()

src = page("source")
h = src[1]
data = src[2:]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: 0 for s in stalls}
for r in data:
    for s in stalls:
        try:
            tot[s] += int(r[ix[s]])
        except ValueError:
            pass
T = sum(tot.values()) or 1
print("stall mix (all samples):", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / T)
                                            for s, v in sorted(tot.items(), key=lambda x: -x[1])[:9]))
print("instructions executed:", sum(int(r[ix["Instructions Executed"]]) for r in data))
print("hottest SASS by samples:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
    print("%6s %9s  %-70s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70],
                                 " ".join("%s=%s" % (s[6:], r[ix[s]]) for s in stalls if int(r[ix[s]]) > 40)))
