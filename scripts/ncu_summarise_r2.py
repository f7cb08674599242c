"""Turn the raw ncu CSV of scripts/gpu/profiles_r2.sh into the committed summaries under profiles/.

    python scripts/ncu_summarise_r2.py r2g
Writes  profiles/<tag>_ncu_launch_summary.csv   per kernel: launches, time, share, DRAM bytes, achieved GB/s
        profiles/<tag>_ncu_launches.csv         every launch of one replayed step
        profiles/hbm_families.json              achieved HBM GB/s per kernel family (bench.py roofline.hbm)
        profiles/gemm_traffic.json              DRAM bytes per GEMM launch (bench.py roofline.traffic)
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6544.7

rows = list(csv.reader(l for l in open(os.path.join(G, f"launches_{tag}.csv")) if l.startswith('"')))
hdr = rows[0]
ii, ik, im, iv, iu = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6}
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(int(r[ii]), {"k": r[ik]})
    d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
ids = sorted(per)
short = lambda n: n.split("(")[0].replace("void ", "").replace("bf::", "")[:72]
# one replayed step = between the last two weight-mirror casts (cast16 over the 28.9 M parameters: > 20 us)
casts = [i for i in ids if "cast16_kernel<__nv_bfloat16>" in per[i]["k"] and per[i].get("gpu__time_duration.sum", 0) > 20000]
s, e = casts[-2], casts[-1]
step = [per[i] for i in ids if s <= i < e]


def family(n):
    if "gemm_tcgen05" in n: return "gemm (tcgen05)"
    if "attn" in n: return "attention"
    if "inorm" in n or "resid_bwd" in n or "colsum16" in n: return "instance-norm / residual / column-sum passes"
    if "patch" in n or "s2d_gather" in n or "cast16" in n or "convert16" in n or "window" in n: return "patch boundary, gathers, casts"
    if "lploss" in n: return "loss"
    if "film" in n or "feat_consts" in n or "branch_param" in n: return "per-channel parameter kernels"
    return "torch glue (fills, copies, rng)"


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in step:
    for tbl, key in ((agg, short(d["k"])), (fam, family(d["k"]))):
        a = tbl[key]
        a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0) / 1e3
        a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{tag}_ncu_launch_summary.csv"), "w") as f:
    f.write("kernel,launches,total_us,share,us_per_launch,dram_read_MB_per_launch,dram_write_MB_per_launch,achieved_GBps,frac_of_hbm_peak\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (a[2] + a[3]) / (a[1] * 1e-6) / 1e9 if a[1] > 0 else 0.0
        f.write(f"\"{k}\",{a[0]},{a[1]:.1f},{a[1] / tot:.4f},{a[1] / a[0]:.1f},{a[2] / a[0] / 1e6:.1f},{a[3] / a[0] / 1e6:.1f},{gbs:.0f},{gbs / PEAK:.3f}\n")
    f.write(f"\"TOTAL (one replayed step, {len(step)} launches, cold cache / serialised)\",{len(step)},{tot:.1f},1.0,,,,,\n")
with open(os.path.join(P, f"{tag}_ncu_launches.csv"), "w") as f:
    f.write("index,kernel,duration_ns,dram_read_bytes,dram_write_bytes\n")
    for j, d in enumerate(step):
        f.write(f"{j},\"{short(d['k'])}\",{d.get('gpu__time_duration.sum', 0):.0f},{d.get('dram__bytes_read.sum', 0):.0f},{d.get('dram__bytes_write.sum', 0):.0f}\n")
out = {"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over the "
                 f"{len(step)} launches of one replayed config-2 step (profiles/{tag}_ncu_launches.csv, scripts/gpu/profiles_r2.sh); "
                 f"achieved = (dram read + write bytes) / device time per family, peak = measured {PEAK} GB/s",
       "peak_GBps": PEAK, "step_us_under_ncu": tot, "families": {}}
for k, a in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / (a[1] * 1e-6) / 1e9 if a[1] > 0 else 0.0
    out["families"][k] = {"launches": a[0], "us": round(a[1], 1), "share_of_step": round(a[1] / tot, 4),
                          "dram_GB": round((a[2] + a[3]) / 1e9, 3), "achieved_GBps": round(gbs, 0), "frac_of_peak": round(gbs / PEAK, 3)}
json.dump(out, open(os.path.join(P, "hbm_families.json"), "w"), indent=1)
g = [d for d in step if "gemm_tcgen05" in d["k"]]
rd, wr = sum(d.get("dram__bytes_read.sum", 0) for d in g), sum(d.get("dram__bytes_write.sum", 0) for d in g)
json.dump({"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the {len(g)} gemm_tcgen05_kernel launches of one "
                     f"replayed config-2 step (profiles/{tag}_ncu_launches.csv; scripts/gpu/profiles_r2.sh)",
           "launches": len(g), "dram_bytes_per_launch": (rd + wr) / max(len(g), 1), "dram_read_bytes_per_step": rd,
           "dram_write_bytes_per_step": wr, "gemm_us_per_step_under_ncu": sum(d.get("gpu__time_duration.sum", 0) for d in g) / 1e3},
          open(os.path.join(P, "gemm_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
