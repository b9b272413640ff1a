"""Digest an ncu --set full report of conv_tc_kernel launches into a text summary for profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("lts__t_bytes.sum", "l2_bytes"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%")]
idx = [(hdr.index(n), lab) for n, lab in want if n in hdr]
print("# " + rep)
print(" | ".join(f"{lab}[{units[i]}]" for i, lab in idx))
for d in data:
    print(" | ".join((d[i][:34] if lab == "kernel" else d[i][:12]) for i, lab in idx))
# stall reasons of the first launch
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::1"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
h = srows[1]
stalls = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(int(r[i] or 0) for r in srows[2:] if len(r) > i) for i, c in stalls}
tot = sum(agg.values()) or 1
print("\n# warp stall sampling, first captured launch (all warps, all roles):")
for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {c:28s} {100 * v / tot:5.1f}%")
