#!/bin/bash
# per-kernel durations (serialised, alone) of the staged filter on outlier-heavy pairs
source tools/r02/gpu_fn.sh
A="--kind sift --images 48 --steps 1 --warmup 1 --no-stages --no-configs --no-cpu-baseline --no-e2e --outlier-frac 0.5"
python bench.py $A > /dev/null 2>&1; echo "plain exit $?"
ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"rs_|fmat_ransac" -c 150 --csv --log-file gpurun_out/r2_staged_launches.csv python bench.py $A > gpurun_out/r2_staged_launches.log 2>&1; echo "ncu exit $?"
python - <<'PYEOF'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_staged_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ii],{"name":r[ki][:40]})[r[mi]]=r[vi]
for k,v in list(d.items())[-44:]:
    print(k, v["name"], v.get("gpu__time_duration.sum"), v.get("launch__grid_size"), v.get("launch__registers_per_thread"), v.get("sm__warps_active.avg.pct_of_peak_sustained_active"), v.get("smsp__issue_active.avg.pct_of_peak_sustained_active"))
PYEOF
