"""Summarise one kernel of an .ncu-rep (ncu --set full) into JSON: the metrics DESIGN.md / bench.py quote, plus the stall mix.
usage: python tools/ncu_summary.py report.ncu-rep [out.json] [note]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "lts__t_sector_op_read_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "sm__cycles_elapsed.avg.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_ltcfabric.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "nvlrx__bytes.sum", "nvltx__bytes.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active"]
out = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    d = {"kernel": r[hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else ""}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            d[w] = [r[i], units[i]]
    stalls = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i] or 0)
    tot = sum(stalls.values()) or 1.0
    d["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.005}
    out.append(d)
if len(sys.argv) > 3:
    for d in out:
        d["_note"] = sys.argv[3]
js = json.dumps(out[0] if len(out) == 1 else out, indent=1)
if len(sys.argv) > 2 and sys.argv[2] != "-":
    open(sys.argv[2], "w").write(js + "\n")
print(js)
