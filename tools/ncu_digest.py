#!/usr/bin/env python
"""Digest of one ncu capture for profiles/: the raw metrics that matter for the roofline block,
the per-opcode totals and the hottest SASS lines (tools/ncu_src_summary.py).
  python tools/ncu_digest.py <stem>_raw.csv <stem>_src.csv "<title>" > profiles/...txt
With --json I L K: also print the ncu_metrics.json record of the first kernel (last line)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
raw, src, title = sys.argv[1], sys.argv[2], sys.argv[3]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
print(title)
recs = []
for v in rows[2:]:
    rec = {}
    print("kernel: %s" % v[hdr.index("Kernel Name")])
    for n in WANT:
        if n in hdr:
            i = hdr.index(n)
            print("%-96s %-16s %s" % (n, units[i], v[i]))
            rec[n] = (units[i], v[i])
    recs.append(rec)
    print()
print("per-opcode totals and hottest lines (tools/ncu_src_summary.py on --page source --csv --print-source sass;"
      " all captured launches together):")
sys.stdout.flush()
subprocess.call([sys.executable, os.path.join(ROOT, "tools", "ncu_src_summary.py"), src, "12"])
if "--json" in sys.argv:
    from bench import csrc_sha16
    k = sys.argv.index("--json")
    I, L, K = (int(x) for x in sys.argv[k + 1:k + 4])
    r = recs[0]
    def val(n, scale=1.0):
        u, x = r[n]
        x = float(x)
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0) * scale
    print(json.dumps({
        "csrc_sha16": csrc_sha16(), "I": I, "L": L, "K": K, "capture": os.path.basename(raw),
        "kernel_ms_under_ncu": val("gpu__time_duration.sum") if r["gpu__time_duration.sum"][0] == "ms" else None,
        "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
        "fp64_pipe_pct": val("TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
        "tensor_pipe_pct": val("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
        "lsu_wavefront_pct": val("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active")}))
