#!/bin/bash
mkdir -p gpurun_out
for f in 0 12288; do
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --debug-flags $f > gpurun_out/bench_f$f.log 2>gpurun_out/bench_f$f.err; echo "bench flags=$f exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_f$f.log").read().strip().splitlines()[-1])
    print("flags $f: value %.0f pairs/s  ms/step %.1f  frac %.3f  knn %.3f ms share %.2f e2e %.0f clocks %s stages %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["share_of_step"], d["e2e"]["value"], d["clocks"], d.get("stages")))
except Exception as e: print("parse fail", e); print(open("gpurun_out/bench_f$f.err").read()[-800:])
PYEOF
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stages"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_i8.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
python - <<'PYEOF'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_i8.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4].split("(")[0][:60]; agg[k][0]+=1; agg[k][1]+=float(r[-1])/1e3
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-62s %4d %9.1f us total %6.3f share %8.1f avg us"%(k,v[0],v[1],v[1]/tot,v[1]/v[0]))
PYEOF
