"""Summarise every kernel launch of an .ncu-rep (ncu --set full) as JSON for profiles/ (profiling aid).
usage: python tools/ncu_to_json.py report.ncu-rep out.json "<command that was profiled>" "<note>" """
import csv, json, subprocess, sys
rep, out, cmd, note = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); h, units = r[0], r[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed.avg.per_cycle_active",
        "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"]
res = []
for v in r[2:]:
    d = {"_kernel": v[h.index("Kernel Name")]}
    for k in KEYS:
        if k in h:
            x = v[h.index(k)]
            try: x = float(x)
            except ValueError: pass
            d[k] = {"value": x, "unit": units[h.index(k)]}
    st = [(float(x), k[len("smsp__pcsamp_warps_issue_stalled_"):]) for k, x in zip(h, v)
          if k.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in k and x not in ("", None)]
    tot = sum(x for x, _ in st) or 1.0
    d["_stall_share_pct"] = {k: round(x / tot * 100, 1) for x, k in sorted(st, reverse=True)[:10]}
    def val(k):
        e = d.get(k); return e["value"] if e else 0.0
    def to_bytes(k):
        e = d.get(k)
        if not e: return 0.0
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(e["unit"], 1)
        return e["value"] * mult
    d["_dram_bytes_per_launch"] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    res.append(d)
json.dump({"_command": cmd, "_note": note, "launches": res}, open(out, "w"), indent=1)
print("wrote", out, len(res), "launches")
