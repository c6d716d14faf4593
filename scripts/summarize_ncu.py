"""Turn ncu outputs in gpurun_out/ into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_red.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return [dict(zip(rows[0], r)) for r in rows[2:]], dict(zip(rows[0], rows[1]))


def launches(path, out_name):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(OUT, out_name), "w") as f:
        f.write("share_pct,launches,median_us,total_us,kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%.2f,%d,%.1f,%.1f,\"%s\"\n" % (100 * sum(v) / tot, len(v), sorted(v)[len(v) // 2] / 1e3, sum(v) / 1e3, k[:110]))


def main():
    os.makedirs(OUT, exist_ok=True)
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    g = os.path.join(ROOT, "gpurun_out")
    for name in sorted(os.listdir(g)):
        p = os.path.join(g, name)
        if name.startswith("final_") and name.endswith(".ncu-rep"):
            recs, units = raw(p)
            base = "%s_%s" % (tag, name[len("final_"):-len(".ncu-rep")])
            with open(os.path.join(OUT, base + ".txt"), "w") as f:
                for rec in recs:
                    f.write("kernel: %s\n" % rec.get("Kernel Name"))
                    f.write("grid %s block %s\n" % (rec.get("Grid Size"), rec.get("Block Size")))
                    for k in KEYS:
                        if k in rec and rec[k] != "":
                            f.write("  %-90s %s %s\n" % (k, rec[k], units.get(k, "")))
                    f.write("\n")
            if name == "final_roi_align_fwd_rows.ncu-rep" and recs:
                rec = recs[0]
                conv = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
                rd = float(rec["dram__bytes_read.sum"]) * conv[units["dram__bytes_read.sum"]]
                wr = float(rec["dram__bytes_write.sum"]) * conv[units["dram__bytes_write.sum"]]
                json.dump({"kernel": rec.get("Kernel Name"), "dram_bytes_read": rd, "dram_bytes_write": wr,
                           "dram_bytes_per_launch": rd + wr, "source": base + ".txt (ncu --set full, one launch)"},
                          open(os.path.join(OUT, "roi_align_fwd_ncu.json"), "w"), indent=1)
        if name.startswith("final_") and name.endswith(".csv"):
            launches(p, "%s_%s" % (tag, name[len("final_"):]))
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
