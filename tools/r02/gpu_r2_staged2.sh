#!/bin/bash
# staged filter, second cut (compact model list, two-step classifier, iterative sampler replay)
source tools/r02/gpu_fn.sh
timeout 1200 python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider -k "fmat or philox or eight_point or pair_body or fountain or staged or essential" > gpurun_out/r2_tests_rs.log 2>&1; echo "ransac tests exit $?"; tail -12 gpurun_out/r2_tests_rs.log
PM_B200_LIB=$PWD/ab/libpm_paranoid.so timeout 900 python tools/ransac_paranoid.py > gpurun_out/r2_staged_paranoid.log 2>&1; echo "paranoid exit $?"
echo "mismatch lines: $(grep -c MISMATCH gpurun_out/r2_staged_paranoid.log)  active: $(grep -c 'paranoid build active' gpurun_out/r2_staged_paranoid.log)"; grep -v "paranoid build active" gpurun_out/r2_staged_paranoid.log | tail -3 | cut -c1-200
A="--kind sift --images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline --no-e2e"
run st2_heavy $A --outlier-frac 0.5
run st2_heavy_b512 $A --outlier-frac 0.5 --batch-pairs 512
run st2_of03 $A --outlier-frac 0.3
run st2_of0 $A
run st2_of0_1k $A --debug-flags 2097152
A="--kind sift --images 48 --steps 1 --warmup 1 --no-stages --no-configs --no-cpu-baseline --no-e2e --outlier-frac 0.5"
ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"rs_|fmat_ransac" -c 150 --csv --log-file gpurun_out/r2_staged_launches.csv python bench.py $A > gpurun_out/r2_staged_launches.log 2>&1; echo "ncu exit $?"
python - <<'PYEOF'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_staged_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ii],{"name":r[ki][:40]})[r[mi]]=r[vi]
for k,v in list(d.items())[107:126]:
    print(k, v["name"], v.get("gpu__time_duration.sum"), v.get("launch__grid_size"), v.get("launch__registers_per_thread"), v.get("sm__warps_active.avg.pct_of_peak_sustained_active"), v.get("smsp__issue_active.avg.pct_of_peak_sustained_active"))
PYEOF
