#!/bin/bash
# SuperPoint (256-d real-valued rows): steady-state bench, launch list, one ncu --set full capture of the MODE 3 kernel
mkdir -p gpurun_out
timeout 900 python bench.py --kind superpoint --images ${SP_IMAGES:-72} --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_superpoint_big.json 2> gpurun_out/bench_superpoint_big.err; echo "bench exit $?"
python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_superpoint_big.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("sp: value %.0f pairs/s ms/step %.1f | knn %.1f frac %.3f avg %.3f ms share %.3f | e2e %.0f | cpu %s | %s %s" % (d["value"], d["ms_per_step"], r["achieved"], r["frac"], r["avg_launch_ms"], r["share_of_step"], d["e2e"]["value"], d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"], r.get("rerank")))
except Exception as e: print("parse fail", e); print(open("gpurun_out/bench_superpoint_big.err").read()[-1500:])
PYEOF
CMD="python bench.py --kind superpoint --images 23 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_launches_sp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sp.csv $CMD > gpurun_out/ncu_launches_sp.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:l2_top2_tc2 -s 1 -c 1 -f -o gpurun_out/prof_tc2f $CMD > gpurun_out/ncu_full_sp.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full_sp.log
ncu --set full --clock-control none --import-source on -k regex:l2f_fixup -s 1 -c 1 -f -o gpurun_out/prof_l2f $CMD > gpurun_out/ncu_full_l2f.log 2>&1
echo "ncu full l2f exit $?"
