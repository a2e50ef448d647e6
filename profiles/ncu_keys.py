#!/usr/bin/env python
"""Print the handful of metrics we judge a kernel by from an `ncu --set full` report.
usage: python profiles/ncu_keys.py gpurun_out/X.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_tensor.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    for path in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
            print(f"## {d.get('Kernel Name', ('?',))[0]}  [{path.split('/')[-1]}]  grid {d.get('Grid Size', ('', ''))[0]} block {d.get('Block Size', ('', ''))[0]}")
            for k in KEYS:
                if k in d and d[k][0] != "":
                    print(f"  {k:82s} {d[k][0]:>16s} {d[k][1]}")
            st = [(float(v[0]), h[len(STALL):]) for h, v in d.items() if h.startswith(STALL) and h.endswith("_per_warp_active.pct") and v[0]]
            st.sort(reverse=True)
            print("  stalls (% of warp-active cycles): " + ", ".join(f"{n.replace('_per_warp_active.pct', '')} {v:.0f}" for v, n in st[:7]))


if __name__ == "__main__":
    main()
