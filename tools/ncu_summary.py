"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the handful of counters DESIGN.md
and profiles/ cite.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [points]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
pts = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    return hdr.index(name) if name in hdr else None
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
keys += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
out = []
for r in data:
    d = {"kernel": r[col("Kernel Name")][:80]}
    for k in keys:
        c = col(k)
        if c is not None:
            try:
                d[k] = float(r[c].replace(",", ""))
            except ValueError:
                d[k] = r[c]
            d[k + "|unit"] = units[c]
    out.append(d)
for d in out:
    print("=" * 100); print(d["kernel"])
    t = d["gpu__time_duration.sum"]; tu = d["gpu__time_duration.sum|unit"]
    tms = t * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[tu]
    def gb(k):
        v, u = d[k], d[k + "|unit"]
        return v * {"Gbyte": 1, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}[u]
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    print(f"time {tms:.3f} ms  dram read {rd:.3f} GB write {wr:.3f} GB  -> {(rd+wr)/tms:.1f} GB/ms = {(rd+wr)/tms*1000:.0f} GB/s")
    if pts:
        print(f"dram bytes per point: {(rd+wr)*1e9/pts:.1f}   instructions per 32 points: {d['smsp__inst_executed.sum']/(pts/32):.0f}")
    for k in keys[3:]:
        if k in d and not k.startswith("smsp__average"):
            print(f"  {k:75s} {d[k]:>16,.2f} {d[k+'|unit']}")
    st = sorted(((d[k], k) for k in keys if k.startswith("smsp__average") and k in d), reverse=True)
    print("  stalls (warps per issue):", ", ".join(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, k in st[:9]))
if len(sys.argv) > 3:
    json.dump(out, open(sys.argv[3], "w"), indent=1)
