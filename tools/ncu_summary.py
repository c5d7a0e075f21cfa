"""Summarise an Nsight Compute report (read on the CPU box with ``ncu -i``)
into the handful of metrics the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/prof.summary.txt
"""

import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]
STALLS = "smsp__average_warps_issue_stalled_"


def page(report, name):
    out = subprocess.run(["ncu", "-i", report, "--page", name, "--csv"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    report = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rows = page(report, "raw")
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(rows[2:]):
        print("== launch {}: {}".format(n, r[ix["Kernel Name"]]))
        for key in RAW:
            if key in ix:
                print("  {:70s} {:>16s} {}".format(key, r[ix[key]],
                                                  units[ix[key]]))
        stalls = sorted(((float(r[i] or 0), h[len(STALLS):].split("_per_")[0])
                         for h, i in ix.items()
                         if h.startswith(STALLS) and h.endswith(
                             "_per_issue_active.ratio")), reverse=True)
        print("  stall cycles per issued instruction: " + ", ".join(
            "{}={:.2f}".format(k, v) for v, k in stalls[:6]))
    src = page(report, "source")
    if len(src) > 2:
        hdr = src[1]
        ix = {h: i for i, h in enumerate(hdr)}
        if "# Samples" in ix:
            data = []
            for r in src[2:]:
                if r and r[0] == "Kernel Name":      # next launch
                    break
                if len(r) == len(hdr):
                    data.append(r)
            total = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
            print("== hottest SASS lines of the first kernel "
                  "(warp-stall samples, {} total)".format(total))
            first, last = ix["stall_barrier"], ix["stall_wait"]
            for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0)
                            )[:top]:
                why = {hdr[i]: int(r[i]) for i in range(first, last + 1)
                       if r[i] and int(r[i]) > 0}
                print("  {:5.1f}%  {:60s} {}".format(
                    100.0 * int(r[ix["# Samples"]]) / total,
                    r[ix["Source"]].strip()[:60], why))


if __name__ == "__main__":
    main()
