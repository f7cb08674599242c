"""Turn the raw ncu CSVs of scripts/ncu_profiles.sh into the committed summaries under profiles/.

    python scripts/ncu_summarise.py r1q
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1q"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def read(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    return hdr, [r for r in rows[hi + 1:] if len(r) == len(hdr)]


def short(name):
    return name.split("(")[0].replace("void ", "")[:70]


# (1) launch list -> one replayed step (between two weight-cast launches), per-kernel totals and shares
hdr, data = read(os.path.join(G, f"launches_{tag}.csv"))
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
names = [r[ik] for r in data]
vals = [float(r[iv].replace(",", "")) for r in data]
starts = [i for i, n in enumerate(names) if "cast16_kernel" in n and vals[i] > 20000]
s, e = starts[-2], starts[-1]          # the last complete step in the capture (a graph replay)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in zip(names[s:e], vals[s:e]):
    agg[short(n)][0] += 1
    agg[short(n)][1] += v / 1000.0
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"{tag}_ncu_launch_summary.csv"), "w") as f:
    f.write("kernel,launches,total_us,share,us_per_launch\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{v[0]},{v[1]:.1f},{v[1] / tot:.4f},{v[1] / v[0]:.1f}\n")
    f.write(f"\"TOTAL (one replayed step, {e - s} launches, cold cache / serialised)\",{e - s},{tot:.1f},1.0,\n")
with open(os.path.join(P, f"{tag}_ncu_launches.csv"), "w") as f:
    f.write("index,kernel,duration_ns\n")
    for i in range(s, e):
        f.write(f"{i - s},\"{short(names[i])}\",{vals[i]:.0f}\n")
print("step launches", e - s, "total ms", tot / 1000)

# (2) DRAM traffic per GEMM launch
hdr, data = read(os.path.join(G, f"gemm_traffic_{tag}.csv"))
ii, ik, im, iv, iu = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = collections.defaultdict(dict)
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
for r in data:
    per[r[ii]]["k"] = short(r[ik])
    per[r[ii]][r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
rd = sum(v.get("dram__bytes_read.sum", 0) for v in per.values())
wr = sum(v.get("dram__bytes_write.sum", 0) for v in per.values())
ns = sum(v.get("gpu__time_duration.sum", 0) for v in per.values())
n = len(per)
out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the {n} gemm_tcgen05_kernel launches of one "
                 f"replayed config-2 step (profiles/{tag}_gemm_traffic.csv; scripts/ncu_profiles.sh)",
       "launches": n, "dram_bytes_per_launch": (rd + wr) / n, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
       "gemm_us_per_step_under_ncu": ns / 1000.0}
json.dump(out, open(os.path.join(P, "gemm_traffic.json"), "w"), indent=1)
with open(os.path.join(P, f"{tag}_gemm_traffic.csv"), "w") as f:
    f.write("launch,kernel,dram_read_bytes,dram_write_bytes,duration_ns\n")
    for k, v in per.items():
        f.write(f"{k},\"{v['k']}\",{v.get('dram__bytes_read.sum', 0):.0f},{v.get('dram__bytes_write.sum', 0):.0f},{v.get('gpu__time_duration.sum', 0):.0f}\n")
print(out)

# (3) full-section capture -> the metrics that matter, per launch
rep = os.path.join(G, f"prof_gemm_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(P, f"{tag}_ncu_gemm_full.csv"), "w") as f:
        f.write(",".join(f"\"{hdr[i]} [{units[i]}]\"" for i in idx) + "\n")
        for r in rows[2:]:
            f.write(",".join(f"\"{r[i][:60]}\"" for i in idx) + "\n")
    print(open(os.path.join(P, f"{tag}_ncu_gemm_full.csv")).read())
