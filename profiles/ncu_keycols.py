"""Key columns of one kernel from `ncu -i X.ncu-rep --page raw --csv`: metric,value lines (what the r0N_ncu_full_*.csv files hold).

  ncu -i gpurun_out/potrf_df.ncu-rep --page raw --csv > /tmp/raw.csv
  python profiles/ncu_keycols.py /tmp/raw.csv potrf_dataflow > profiles/r02_ncu_full_potrf_dataflow_n4096_v2.csv
"""
import csv
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    pat = sys.argv[2]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h, units = rows[hi], rows[hi + 1]
    kn = h.index("Kernel Name")
    for r in rows[hi + 2:]:
        if len(r) <= kn or pat not in r[kn]:
            continue
        print("metric,value,unit")
        for i, name in enumerate(h):
            if name in KEEP or name.startswith("smsp__average_warps_issue_stalled") and name.endswith("per_issue_active.ratio"):
                print("%s,%s,%s" % (name, r[i].replace(",", ""), units[i] if i < len(units) else ""))
        break


if __name__ == "__main__":
    main()
