"""Summarise `ncu --page raw --csv` exports (one row per profiled launch) into the few numbers the roofline needs.

    python tools/ncu_summary.py gpurun_out/<tag>/full_*.raw.csv > profiles/<name>.txt
"""
import csv
import sys

WANT = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg"]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print("== %s :: %s" % (path.split("/")[-1], r[hdr.index("Kernel Name")][:110]))
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print("   %-70s %s %s" % (w, r[i], units[i]))
            st = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]:
                    st.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            print("   stalls (warps per issue): " + ", ".join("%s=%.2f" % (h, v) for v, h in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    main()
