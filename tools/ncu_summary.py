#!/usr/bin/env python
"""Summarise ncu artefacts brought back from the GPU box into small text files under profiles/.

  python tools/ncu_summary.py list  gpurun_out/launches.csv  profiles/rNN_launch_list.txt  "<command>"
  python tools/ncu_summary.py full  gpurun_out/prof_x.ncu-rep profiles/rNN_x_full.txt      "<command>"
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_uniform.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def read_csv_text(text):
    rows = list(csv.reader(io.StringIO(text)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[start], rows[start + 1:]


def do_list(src, dst, cmd):
    text = open(src).read()
    hdr, rows = read_csv_text(text)
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows:
        if len(r) <= mv or not r[0].isdigit():
            continue
        name = r[kn].split("(")[0][:90]
        ns = float(r[mv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    unit = rows[0][hdr.index("Metric Unit")] if rows else "ns"
    div = {"ns": 1e3, "us": 1.0, "ms": 1e-3, "nsecond": 1e3, "usecond": 1.0, "msecond": 1e-3}.get(unit, 1e3)
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none  {cmd}\n")
        f.write("(per-launch times under ncu are cold-cache and serialised: compare SHARES)\n")
        f.write("kernel, launches, total_us, share\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name}, {n}, {t / div:.1f}, {t / total:.3f}\n")
    print(open(dst).read())


def do_full(src, dst, cmd):
    text = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    hdr, rows = read_csv_text(text)
    units, rows = rows[0], rows[1:]
    kn = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"ncu --set full --clock-control none --import-source on  {cmd}\nsource report: {src}\n")
        for r in rows:
            f.write(f"\n== launch {r[0]}: {r[kn][:100]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k} = {r[i]} {units[i]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    (do_list if mode == "list" else do_full)(src, dst, cmd)
