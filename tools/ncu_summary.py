"""Key counters of one kernel from an ncu report (raw page) as JSON.
usage: python tools/ncu_summary.py rep.ncu-rep "<note>" > profiles/xyz.json"""
import csv, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.max"]
m = {k: [v[h.index(k)], units[h.index(k)]] for k in want if k in h}
name = v[h.index("Kernel Name")] if "Kernel Name" in h else ""
print(json.dumps({"kernel": name, "note": sys.argv[2] if len(sys.argv) > 2 else "", "metrics": m}, indent=1))
