#!/bin/bash
# kernel names + durations of the GEMM launches of one replayed step, for both GEMM modes (proves which template ran)
TAG=${1:-r2d}
for cg in 1 2; do
  BF_GEMM_CG=$cg ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:gemm_tcgen05 --launch-skip 234 --launch-count 234 --csv \
    --log-file gpurun_out/${TAG}_gemm_cg$cg.csv python bench.py --profile --steps 2 > gpurun_out/${TAG}_ncu_cg$cg.log 2>&1
  echo "cg=$cg rc=$?"
done
python - <<'PY'
import csv, collections, sys
for cg in (1, 2):
    rows = list(csv.reader(l for l in open(f"gpurun_out/%s_gemm_cg{cg}.csv" % sys.argv[1] if len(sys.argv) > 1 else f"gpurun_out/r2d_gemm_cg{cg}.csv") if l.startswith('"')))
    h = rows[0]; ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    iid = h.index("ID")
    per = collections.defaultdict(dict)
    names = {}
    for r in rows[1:]:
        per[r[iid]][r[im]] = float(r[iv].replace(",", ""))
        names[r[iid]] = r[ik]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for k, m in per.items():
        a = agg[names[k][:60]]
        a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0) / 1e3
        a[2] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)
        a[3] += m.get("lts__t_bytes.sum", 0) / 1e6
        a[4] += m.get("sm__cycles_elapsed.avg.per_second", 0) / 1e6
    print(f"== BF_GEMM_CG={cg}")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:62s} n={a[0]:3d} total {a[1]:8.1f} us  avg {a[1]/a[0]:7.1f} us  tensor {a[2]/a[0]:5.1f}%  L2 {a[3]/a[0]:7.1f} MB/launch  clk {a[4]/a[0]:6.0f} MHz")
PY
