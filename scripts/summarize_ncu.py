"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep, scratch) into the tracked text file the DESIGN cites and
refresh profiles/traffic.json (dram bytes per launch of the dense kernels bench.py reports as roofline.traffic).
usage: python scripts/summarize_ncu.py OUT.txt REPORT.ncu-rep [REPORT.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpc__cycles_elapsed.max.per_second"]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
KERNEL_ID = {"dense_scan_kernel": "1", "dense_tc_kernel": "2", "dense_tc2_kernel": "3", "dense_tc2cvt_kernel": "5"}


def main():
    out_path, reports = sys.argv[1], sys.argv[2:]
    lines = ["# ncu --set full --clock-control none --import-source on; bench.py c3 (10M x 768 fp32 + postings nnz 8.8e8), batch 256 (K1: batch 1)",
             "# command: scripts/gpu_final.sh; reports (scratch): " + " ".join(reports), ""]
    traffic = {}
    for rep in reports:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            lines.append(f"## {rep}: no kernel captured\n")
            continue
        head, units = rows[0], rows[1]
        ki = head.index("Kernel Name")
        for r in rows[2:]:
            lines.append("## " + r[ki][:120])
            vals = {}
            for m in METRICS:
                if m in head:
                    i = head.index(m)
                    lines.append(f"{m:<78s} {r[i]:>14s} {units[i]}")
                    vals[m] = (float(r[i].replace(",", "")), units[i])
            for name, kid in KERNEL_ID.items():
                if (name + "<") in r[ki] or (name + "(") in r[ki]:
                    rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
                    if rd and wr:
                        traffic[kid] = rd[0] * UNIT_SCALE.get(rd[1], 1.0) + wr[0] * UNIT_SCALE.get(wr[1], 1.0)
            lines.append("")
    open(out_path, "w").write("\n".join(lines))
    # merge: kernels that are not in these reports keep the traffic of the capture they were last seen in
    try:
        tj = json.load(open("profiles/traffic.json"))
    except Exception:
        tj = {"_source": "", "c3": {}}
    if traffic:
        tj["c3"].update(traffic)
        tj["_source"] = f"ids {sorted(traffic)}: {out_path}; others: " + tj.get("_source", "")[:300]
        json.dump(tj, open("profiles/traffic.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
