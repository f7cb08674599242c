"""Summarise one `ncu --set full --import-source on` capture of a single kernel launch (scripts/gpu/ncu_one.sh) into text.

    python scripts/ncu_kernel_summary.py gpurun_out/prof_<name>_<tag>.ncu-rep > profiles/<tag>_ncu_<name>.txt
Prints the launch's key metrics (duration, DRAM bytes and throughput, tensor pipe, issue rate, occupancy, registers, shared
memory), the warp stall distribution and the SASS instructions with the most stall samples.
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
print("capture:", rep)
print("kernel :", m.get("Kernel Name", ("?", ""))[0][:110])
keys = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (occupancy)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / block"),
    ("launch__occupancy_limit_shared_mem", "blocks / SM (shared memory limit)"),
    ("launch__occupancy_limit_registers", "blocks / SM (register limit)"),
]
for k, name in keys:
    if k in m and m[k][0] not in ("", "n/a"):
        print(f"  {name:34s} {m[k][0]} {m[k][1]}")
try:
    t_us = float(m["gpu__time_duration.sum"][0].replace(",", ""))
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    by = sum(float(m[k][0].replace(",", "")) * scale.get(m[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    tu = {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3}.get(m["gpu__time_duration.sum"][1], 1.0)
    print(f"  {'DRAM bytes / time':34s} {by / (t_us * tu) / 1e3:.0f} GB/s")
except Exception:
    pass
stall = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(v) for h, v in zip(hdr, vals)
         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")}
tot = sum(stall.values()) or 1.0
print("warp stall reasons (share of warp cycles per issued instruction):")
for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {k:24s} {100 * v / tot:5.1f} %")

src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
if starts:
    s = starts[0]
    h = src[s + 1]
    data = [r for r in src[s + 2:] if len(r) == len(h)]
    ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    total = sum(int(r[isamp]) for r in data) or 1
    print(f"SASS: {len(data)} instructions, {total} stall samples; instructions with the most samples:")
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:ntop]):
        r = data[i]
        why = {h[j][6:]: int(r[j]) for j in st if int(r[j]) > 0}
        why = ", ".join(f"{k} {v}" for k, v in sorted(why.items(), key=lambda kv: -kv[1])[:2])
        print(f"  #{i:5d} {100 * int(r[isamp]) / total:5.1f} %  x{r[iex]:>8s}  {r[ia].strip()[:64]:64s} {why}")
