"""One-line roofline summary per kernel of an .ncu-rep (ncu --set full): python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]
Prints CSV: kernel, grid, regs, duration_us, dram_read_MB, dram_write_MB, dram_pct_of_peak, tensor_pipe_pct, xu_pipe_pct,
issue_active_pct, warps_active_pct, l2_hit_pct, shared_bank_conflicts."""
import csv, subprocess, sys
KEYS = [("gpu__time_duration.sum", "duration_us"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "xu_pct"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
        ("smsp__inst_executed.sum", "warp_instructions")]
print(",".join(["kernel"] + [k[1] for k in KEYS]))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r))
        u = dict(zip(h, units))
        name = d.get("Kernel Name", "?").replace("<unnamed>::", "").replace("__nv_bfloat16", "bf16")
        name = name.split("(")[0][:70]
        vals = []
        for k, _ in KEYS:
            v = d.get(k, "")
            try:
                f = float(v)
                if u.get(k) == "byte": f /= 1e6
                if u.get(k) == "Gbyte": f *= 1e3
                if u.get(k) == "Kbyte": f /= 1e3
                if u.get(k) == "ns": f /= 1e3
                if u.get(k) == "ms": f *= 1e3
                v = "%.4g" % f
            except ValueError:
                pass
            vals.append(v)
        print(",".join(['"%s"' % name] + vals))
