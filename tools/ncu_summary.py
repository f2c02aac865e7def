"""Summarise .ncu-rep files (read here, no GPU): key throughput metrics + top stall reasons.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_shared_ld.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print(f"== {path}: {d.get('Kernel Name', ('?',))[0]}  grid={d.get('Grid Size', ('?',))[0]} block={d.get('Block Size', ('?',))[0]}")
        for k in KEYS:
            if k in d:
                print(f"   {k:78s} {d[k][0]:>16s} {d[k][1]}")
        stalls = [(h, float(v.replace(',', ''))) for h, (v, u) in d.items()
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and v]
        stalls.sort(key=lambda t: -t[1])
        for h, v in stalls[:6]:
            print(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.2f}")


if __name__ == "__main__":
    main()
