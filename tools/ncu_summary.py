"""Summarise an .ncu-rep (one kernel launch per row) into the text that is committed under
profiles/.  usage: python tools/ncu_summary.py report.ncu-rep > profiles/NAME.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of %s (ncu --set full --clock-control none; times under the profiler are" % rep.split("/")[-1])
    print("# cold-cache and serialised -- they are evidence for traffic / occupancy / mix, not bench values)")
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("\nkernel: %s" % r[ik][:160])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-78s %s %s" % (k, r[i], units[i]))
        if "dram__bytes_read.sum" in hdr and "gpu__time_duration.sum" in hdr:
            def val(k):
                i = hdr.index(k)
                v = float(r[i].replace(",", ""))
                u = units[i].lower()
                scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6,
                         "ms": 1e-3, "s": 1.0}.get(u, 1.0)
                return v * scale
            tr = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            t = val("gpu__time_duration.sum")
            print("  %-78s %.6e byte" % ("derived: dram traffic per launch (read + write)", tr))
            print("  %-78s %.1f GB/s" % ("derived: dram traffic / duration", tr / t / 1e9))


if __name__ == "__main__":
    main()
