"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv):
per-kernel launches / time / DRAM bytes, and the per-launch-per-candidate DRAM traffic of the DMMA GEMM that
bench.py reports as roofline.traffic.

  python profiles/ncu_traffic.py profiles/r02_launches_one_step_b32.csv 32 > profiles/r02_gemm_traffic.json
"""
import collections
import csv
import json
import sys


def main():
    path, batch = sys.argv[1], int(sys.argv[2])
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    kn, mn, mv, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.defaultdict(dict)
    names = {}
    for r in rows[hi + 1:]:
        if len(r) <= mv or not r[idc].isdigit():
            continue
        per[int(r[idc])][r[mn]] = float(r[mv].replace(",", ""))
        names[int(r[idc])] = r[kn]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for i, m in per.items():
        k = names[i].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg[k]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in agg.values())
    gemm = [a for k, a in agg.items() if k.startswith("gemm_nt")]
    glaunch, gbytes = sum(a[0] for a in gemm), sum(a[2] for a in gemm)
    out = {"source": path, "batch": batch, "launches": sum(a[0] for a in agg.values()),
           "dram_bytes_per_gemm_launch_per_candidate": gbytes / max(glaunch, 1) / batch,
           "gemm_launches": glaunch, "gemm_dram_GB": gbytes / 1e9,
           "kernels": {k: {"launches": a[0], "ms": a[1] / 1e6, "share_of_kernel_time": a[1] / total, "dram_GB": a[2] / 1e9}
                       for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
