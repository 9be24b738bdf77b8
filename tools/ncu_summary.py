"""Summarise an .ncu-rep (`ncu --set full`) into the few metrics the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [> profiles/rNN_ncu_X_summary.txt]

Reads the raw page through `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("%-70s %s" % ("Kernel Name", d.get("Kernel Name")))
        for k in KEYS:
            if k in d:
                print("%-70s %s %s" % (k, d[k], u.get(k, "")))
        stalls = []
        for k, v in d.items():
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
                try:
                    stalls.append((float(v.replace(",", "")), k[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print("warp stall reasons (pc sampling):")
        for v, k in sorted(stalls, reverse=True)[:9]:
            print("   %-28s %5.1f%%" % (k, 100 * v / tot))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
